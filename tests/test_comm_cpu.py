"""CPU: the host side of the library's own data-parallel exchange (ssdgeom/comm.py over ssdg_comm_*): the id
broadcast channel, the binding, and the byte-compiled reference build the CPU bench arm runs (oracle/build_ref.py).
The collective itself needs GPUs: tests/test_gpu_round2.py."""
import multiprocessing as mp
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bcast_worker(rank, world, port, q):
    from ssdgeom import comm
    payload = bytes(range(128)) if rank == 0 else None
    q.put((rank, comm.tcp_broadcast(payload, world, rank, "127.0.0.1", port)))


def test_tcp_broadcast_three_ranks():
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(got[r] == bytes(range(128)) for r in range(world))


def test_comm_binding_fails_loudly_without_devices():
    import build_native
    build_native.build()
    from ssdgeom import _native as N, comm
    # NCCL is bound at run time: present in this image, and the argument checks need no GPU
    assert isinstance(comm.available(), bool)
    assert N.lib().ssdg_comm_allreduce_sum(None, None, 1, N.F64, None) == N.ERR_ARG
    assert N.lib().ssdg_comm_init_rank(None, None, 2, 0) == N.ERR_ARG
    with pytest.raises(ValueError):
        comm.Comm(b"short", 1, 0)
    assert "NCCL" in N.status_string(N.ERR_NO_NCCL) or "nccl" in N.status_string(N.ERR_NO_NCCL)


def test_byte_compiled_reference_runs_without_its_source_tree():
    """oracle/build_ref.py: the reference's own modules, py_compile'd from /root/reference into oracle/_ref, import
    source-less (what the GPU box sees) and agree with the oracle port."""
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    out = build_ref.build()
    if out is None:
        pytest.skip("neither /root/reference nor oracle/_ref is present")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from oracle import ref_loader, ssd_oracle as O\n"
        "ns = ref_loader.load(); assert ns.root.endswith('_ref'), ns.root\n"
        "from ssdgeom import synth\n"
        "pri = ns.build_prior_box(synth.SSD300['sizes']); assert np.array_equal(pri, O.build_prior_box())\n"
        "b, c, o = synth.make_gt(5, 1, 12, 'max')\n"
        "lab, box, mask = ns.match_bbox(c, b, pri, 0.5)\n"
        "w = O.match_bbox(c, b, pri, 0.5)\n"
        "assert np.array_equal(lab, w[0]) and np.array_equal(box, w[1]) and np.array_equal(mask, w[2])\n"
        "print('ok')\n" % (ROOT, os.path.join(ROOT, "ssd-object-detection_b200")))
    env = dict(os.environ, SSDGEOM_REFERENCE_ROOT="/nonexistent")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-2000:]
