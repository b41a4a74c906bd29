"""Run bench.py once per environment setting and print value / step / kernel times side by side.
usage: python tools/sweep_env.py [--args "bench args"] "A=1 B=2" "A=3" ...   ("" = defaults)"""
import json, os, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
argv = sys.argv[1:]
extra = []
if argv and argv[0] == "--args":
    extra = argv[1].split(); argv = argv[2:]
for setting in argv:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("=", 1)
        env[k] = v
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline", "--no-e2e", "--no-detail"] + extra,
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        k = d["detail"]["kernels_ms"]
        print("%-44s %9.0f img/s  step %.4f ms  frac %.3f  filter %.4f nms %.4f search %.4f match %.4f tail %.4f" % (
            setting or "(defaults)", d["value"], d["ms_per_step"], d["roofline"]["frac"], k["filter_kernel_with_row_stats"],
            k["nms_kernel"], k["search_kernel"], k["match_kernel"], k["loss_tail"]), flush=True)
    except Exception as e:   # noqa: BLE001
        print(setting, "FAILED", e, p.stderr[-600:], flush=True)
