// Post-processing: softmax score head (models/ssd_model.py:479-488), box decode (:466-467) and the
// per-class score-threshold / top-k / greedy NMS the reference lacks (spec: oracle/ssd_oracle.py
// nms_per_class, IoU formula utils/bbox.py:13-25 in float32).
//
//   filter_kernel  one streaming pass over the logits [N,C]: same per-warp bulk-TMA tile ring as the
//                  loss; lane r owns row r.  Emits (score key, prior) candidates per (image, class)
//                  for p > score_thresh, decodes every box once, optional head outputs / softmax.
//   emit_kernel    the same candidate emission from caller-supplied probabilities (ssdg_nms).
//   nms_kernel     one CTA per (image, class): exact top-k by (score desc, prior asc) -- radix select
//                  when the list is longer than the sort width, then a bitonic sort -- lower-triangle
//                  suppression bit matrix built with warp ballots (division-free margin test, exact
//                  IEEE division only inside the margin), and a parallel fixed-point resolution of the
//                  greedy recurrence "kept(i) = no kept j < i suppresses i".
#include <math_constants.h>
#include "common.cuh"

namespace ssdg {

constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFStages = 2;
constexpr int kNmsThreads = 256;

struct DetectParams {
  const float* pred_cls;   // logits (filter) or probabilities (emit)
  const float* pred_box;
  const void* priors;
  long long N;             // B*A
  int A, C;
  float score_thresh;
  u32* ccount;             // [B*(C-1)]
  u64* cand;               // [B*(C-1)][A]
  float* boxes;            // [N,4] decoded
  float* probs;            // optional [N,C]
  float head_thresh;
  float* head_score;
  int* head_cls;
  uint8_t* head_mask;
};

__device__ __forceinline__ u64 evict_first_policy() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_hint(void* smem_dst, const void* gsrc, u32 bytes, u64* bar, u64 pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

__device__ __forceinline__ void emit_candidate(const DetectParams& P, int b, int a, int c, float score) {
  const size_t list = (size_t)b * (P.C - 1) + c;
  const u32 pos = atomicAdd(&P.ccount[list], 1u);
  if (pos < (u32)P.A) P.cand[list * P.A + pos] = ((u64)key32(score) << 32) | (u64)(~(u32)a);
}

template <typename TP>
__device__ __forceinline__ float4 decode_row(float4 t, const void* priors, int a) {
  double dx, dy, dw, dh;
  if (sizeof(TP) == 8) {
    const double2* p = reinterpret_cast<const double2*>(priors) + 2 * (size_t)a;
    double2 u = __ldg(p), v = __ldg(p + 1);
    dx = u.x; dy = u.y; dw = v.x; dh = v.y;
  } else {
    float4 v = __ldg(reinterpret_cast<const float4*>(priors) + a);
    dx = v.x; dy = v.y; dw = v.z; dh = v.w;
  }
  float4 o;
  o.x = (float)((double)t.x * dw + dx);
  o.y = (float)((double)t.y * dh + dy);
  o.z = (float)(exp((double)t.z) * dw);
  o.w = (float)(exp((double)t.w) * dh);
  return o;
}

// One warp tile (32 priors, lane r owns row r in shared memory; the rows may be overwritten with the
// probabilities).  Candidate emission is warp-cooperative: a ballot per class gives the tile's
// candidate mask, the lanes then reserve list space for all classes at once (one atomic per
// (tile, class), issued side by side so their latencies overlap) and finally scatter their entries.
template <typename TP>
__device__ __forceinline__ void filter_tile(const DetectParams& P, long long n, int b, int a, bool valid, float* row,
                                            u32* wmask, u32* wbase, int lane) {
  const int C = P.C, nfg = P.C - 1;
  float m = 0.f, s = 1.f;
  if (valid) {
    float m0 = -CUDART_INF_F, m1 = m0, m2 = m0, m3 = m0;
    int c = 0;
    for (; c + 4 <= C; c += 4) {
      m0 = fmaxf(m0, row[c]); m1 = fmaxf(m1, row[c + 1]); m2 = fmaxf(m2, row[c + 2]); m3 = fmaxf(m3, row[c + 3]);
    }
    for (; c < C; ++c) m0 = fmaxf(m0, row[c]);
    m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    c = 0;
    for (; c + 4 <= C; c += 4) {
      s0 += __expf(row[c] - m); s1 += __expf(row[c + 1] - m); s2 += __expf(row[c + 2] - m); s3 += __expf(row[c + 3] - m);
    }
    for (; c < C; ++c) s0 += __expf(row[c] - m);
    s = (s0 + s1) + (s2 + s3);
  }
  // p_c > thresh  <=>  x_c - m > log(thresh * s); pre-filter in logit space with a margin, then the
  // exact score  exp(x_c - m) / s  decides.
  const float cut = (P.score_thresh > 0.f) ? __logf(P.score_thresh * s) - 1e-3f : -CUDART_INF_F;
  for (int c = 0; c < nfg; ++c) {
    bool p = false;
    if (valid) {
      const float d = row[c] - m;
      if (d > cut) p = __fdiv_rn(__expf(d), s) > P.score_thresh;
    }
    const u32 mc = __ballot_sync(SSDG_FULL, p);
    if (lane == 0) wmask[c] = mc;
  }
  __syncwarp();
  const int b0 = __shfl_sync(SSDG_FULL, b, 0);
  const u32 other = __ballot_sync(SSDG_FULL, valid && b > b0 + 1);
  if (other) {
    // a tile spanning more than two images (fewer than 32 priors per image): plain per-candidate path
    if (valid)
      for (int c = 0; c < nfg; ++c)
        if ((wmask[c] >> lane) & 1u) emit_candidate(P, b, a, c, __fdiv_rn(__expf(row[c] - m), s));
  } else {
    for (int which = 0; which < 2; ++which) {
      const u32 sel = __ballot_sync(SSDG_FULL, valid && b == b0 + which);
      if (!sel) continue;
      const size_t lbase = (size_t)(b0 + which) * nfg;
      for (int c = lane; c < nfg; c += 32) {
        const int cnt = __popc(wmask[c] & sel);
        wbase[c] = cnt ? atomicAdd(&P.ccount[lbase + c], (u32)cnt) : 0u;
      }
      __syncwarp();
      const u32 lt = (1u << lane) - 1u;
      for (int c = 0; c < nfg; ++c) {
        const u32 mc = wmask[c] & sel;
        if (!mc) continue;
        if ((mc >> lane) & 1u) {
          const float score = __fdiv_rn(__expf(row[c] - m), s);
          const u32 pos = wbase[c] + (u32)__popc(mc & lt);
          if (pos < (u32)P.A) P.cand[(lbase + c) * P.A + pos] = ((u64)key32(score) << 32) | (u64)(~(u32)a);
        }
      }
      __syncwarp();
    }
  }
  if (!valid) return;
  if (P.head_score || P.head_cls || P.head_mask) {
    // models/ssd_model.py:481-488: max foreground probability, arg-max over all classes (first max)
    float best = row[0];
    int arg = 0;
    float fg = -CUDART_INF_F;
    for (int c = 0; c < C; ++c) {
      const float v = row[c];
      if (v > best) { best = v; arg = c; }
      if (c < C - 1) fg = fmaxf(fg, v);
    }
    const float score = __fdiv_rn(__expf(fg - m), s);
    const float pbg = __fdiv_rn(__expf(row[C - 1] - m), s);
    if (P.head_score) P.head_score[n] = score;
    if (P.head_cls) P.head_cls[n] = arg;
    if (P.head_mask) P.head_mask[n] = (score > P.head_thresh && !(pbg > P.head_thresh)) ? 1 : 0;
  }
  if (P.boxes) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(P.pred_box) + n);
    reinterpret_cast<float4*>(P.boxes)[n] = decode_row<TP>(t, P.priors, a);
  }
  if (P.probs)
    for (int c = 0; c < C; ++c) row[c] = __fdiv_rn(__expf(row[c] - m), s);
}

template <typename TP>
__global__ void __launch_bounds__(kFThreads, 1) filter_kernel(DetectParams P, int warps_per_cta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = P.C, A = P.A;
  const u32 tile_bytes = 32u * (u32)C * 4u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* bufs = reinterpret_cast<float*>(smem_raw);
  u64* bars = reinterpret_cast<u64*>(smem_raw + (size_t)warps_per_cta * kFStages * tile_bytes);
  u32* scratch = reinterpret_cast<u32*>(bars + kFWarps * kFStages);   // [warps][2][C]
  if (tid == 0) {
    for (int i = 0; i < warps_per_cta * kFStages; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (warp >= warps_per_cta) return;
  u32* wmask = scratch + (size_t)warp * 2 * C;
  u32* wbase = wmask + C;
  const long long full_tiles = P.N >> 5;
  const long long gw = (long long)blockIdx.x * warps_per_cta + warp;
  const long long stride = (long long)gridDim.x * warps_per_cta;
  float* mybuf = bufs + (size_t)warp * kFStages * 32 * C;
  u64* mybar = bars + warp * kFStages;
  const u64 pol = evict_first_policy();
  const char* src = reinterpret_cast<const char*>(P.pred_cls);
  if (lane == 0) {
    for (int s = 0; s < kFStages; ++s) {
      long long t = gw + (long long)s * stride;
      if (t < full_tiles) {
        mbar_arrive_expect_tx(&mybar[s], tile_bytes);
        tma_load_hint(mybuf + (size_t)s * 32 * C, src + (size_t)t * tile_bytes, tile_bytes, &mybar[s], pol);
      }
    }
  }
  int k = 0;
  for (long long t = gw; t < full_tiles; t += stride, ++k) {
    const int s = k % kFStages;
    mbar_wait(&mybar[s], (u32)((k / kFStages) & 1));
    float* tile = mybuf + (size_t)s * 32 * C;
    const long long n = (t << 5) + lane;
    const int b = (int)(n / A), a = (int)(n - (long long)b * A);
    filter_tile<TP>(P, n, b, a, true, tile + (size_t)lane * C, wmask, wbase, lane);
    __syncwarp();
    if (P.probs) {  // the tile layout in shared memory equals the layout in global memory
      float4* dst = reinterpret_cast<float4*>(P.probs + (size_t)t * 32 * C);
      const float4* s4 = reinterpret_cast<const float4*>(tile);
      for (int i = lane; i < 8 * C; i += 32) __stcs(&dst[i], s4[i]);
      __syncwarp();
    }
    const long long tn = t + (long long)kFStages * stride;
    if (lane == 0 && tn < full_tiles) {
      mbar_arrive_expect_tx(&mybar[s], tile_bytes);
      tma_load_hint(tile, src + (size_t)tn * tile_bytes, tile_bytes, &mybar[s], pol);
    }
  }
  const int tail = (int)(P.N & 31);
  if (gw == 0 && tail) {
    const float* g = P.pred_cls + (size_t)full_tiles * 32 * C;
    for (int i = lane; i < tail * C; i += 32) mybuf[i] = g[i];
    __syncwarp();
    const bool valid = lane < tail;
    const long long n = (full_tiles << 5) + (valid ? lane : 0);
    const int b = (int)(n / A), a = (int)(n - (long long)b * A);
    filter_tile<TP>(P, n, b, a, valid, mybuf + (size_t)(valid ? lane : 0) * C, wmask, wbase, lane);
    __syncwarp();
    if (P.probs)
      for (int i = lane; i < tail * C; i += 32) P.probs[(size_t)full_tiles * 32 * C + i] = mybuf[i];
  }
}

// Candidate emission from given probabilities: thread per (prior, class) element, coalesced.
__global__ void __launch_bounds__(256) emit_kernel(DetectParams P) {
  const long long total = P.N * P.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float p = P.pred_cls[i];
    const long long n = i / P.C;
    const int c = (int)(i - n * P.C);
    if (c < P.C - 1 && p > P.score_thresh) {
      const int b = (int)(n / P.A);
      emit_candidate(P, b, (int)(n - (long long)b * P.A), c, p);
    }
  }
}

// ---- per-(image, class) NMS ---------------------------------------------------------------------------
struct NmsParams {
  const u32* ccount;
  const u64* cand;
  const float* boxes;  // [B,A,4]
  int A, n_fg, top_k, sortn;  // sortn: power of two >= top_k
  float iou_thresh;
  int* out_kept;
  int* out_count;
  float* out_score;
};

__device__ __forceinline__ void bitonic_sort_desc(u64* keys, int n, int tid, int nthreads) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < (n >> 1); i += nthreads) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kNmsThreads) nms_kernel(NmsParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sortn = P.sortn, W = sortn >> 5;
  u64* keys = reinterpret_cast<u64*>(smem_raw);                 // [sortn]
  float4* crn = reinterpret_cast<float4*>(keys + sortn);        // [sortn] x1,y1,x2,y2
  float* area = reinterpret_cast<float*>(crn + sortn);          // [sortn]
  u32* sup = reinterpret_cast<u32*>(area + 2 * sortn);          // [sortn][W] lower triangle (after area, qa)
  u32* keptw = sup + (size_t)sortn * W;                         // [W]
  u32* remw = keptw + W;                                        // [W]
  u32* hist = remw + W;                                         // [256]
  __shared__ u64 sel_prefix;
  __shared__ int sel_k, sel_fill;

  const size_t list = blockIdx.x;
  const int b = (int)(list / P.n_fg);
  int n = (int)min(P.ccount[list], (u32)P.A);
  const u64* cl = P.cand + list * (size_t)P.A;

  for (int i = tid; i < sortn; i += kNmsThreads) keys[i] = 0ull;
  __syncthreads();
  if (n <= sortn) {
    for (int i = tid; i < n; i += kNmsThreads) keys[i] = cl[i];
  } else {
    // Radix select of the top_k-th largest composite key (keys are unique), 8 bits per pass.
    if (tid == 0) { sel_prefix = 0ull; sel_k = P.top_k; sel_fill = 0; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += kNmsThreads) hist[i] = 0u;
      __syncthreads();
      const u64 pre = sel_prefix;
      const u64 hmask = shift == 56 ? 0ull : (~0ull << (shift + 8));
      for (int i = tid; i < n; i += kNmsThreads) {
        const u64 v = cl[i];
        if ((v & hmask) == pre) atomicAdd(&hist[(u32)(v >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int k = sel_k, acc = 0, d = 255;
        for (; d > 0; --d) {
          if (acc + (int)hist[d] >= k) break;
          acc += (int)hist[d];
        }
        sel_k = k - acc;
        sel_prefix = pre | ((u64)d << shift);
      }
      __syncthreads();
    }
    const u64 kth = sel_prefix;
    for (int i = tid; i < n; i += kNmsThreads) {
      const u64 v = cl[i];
      if (v >= kth) {
        const int p = atomicAdd(&sel_fill, 1);
        if (p < sortn) keys[p] = v;
      }
    }
    __syncthreads();
    n = min(sel_fill, sortn);
  }
  int sn = 32;
  while (sn < n) sn <<= 1;           // n <= sortn here; the keys beyond n are 0 and sort to the end
  bitonic_sort_desc(keys, sn, tid, kNmsThreads);
  const int m = min(n, P.top_k);

  // gather decoded boxes; corners in the formula's own float32 operations (utils/bbox.py:13-21).
  // qa = q*(area + 0.5e-10), q = thr/(1+thr):  iou > thr  <=>  inter > qa_i + qa_j  in exact arithmetic
  // (denominator positive); boxes that can never overlap anything (w <= 0, h <= 0, non-finite) get
  // qa = +inf so the fast test rejects them exactly like the formula does (their intersection is 0).
  const float thr = P.iou_thresh;
  const bool fast_ok = thr > 0.f && thr < 1e6f;
  const float q = fast_ok ? thr / (1.f + thr) : 0.f;
  float* qa = area + sortn;   // [sortn]
  for (int i = tid; i < m; i += kNmsThreads) {
    const int a = (int)(~(u32)keys[i]);
    const float4 bx = __ldg(reinterpret_cast<const float4*>(P.boxes) + (size_t)b * P.A + a);
    const float hw = __fmul_rn(bx.z, 0.5f), hh = __fmul_rn(bx.w, 0.5f);
    const float4 cr = make_float4(__fsub_rn(bx.x, hw), __fsub_rn(bx.y, hh), __fadd_rn(bx.x, hw), __fadd_rn(bx.y, hh));
    const float ar = __fmul_rn(bx.z, bx.w);
    crn[i] = cr;
    area[i] = ar;
    const bool sane = bx.z > 0.f && bx.w > 0.f && isfinite(cr.x) && isfinite(cr.y) && isfinite(cr.z) &&
                      isfinite(cr.w) && isfinite(ar);
    qa[i] = sane ? q * (ar + 0.5e-10f) : CUDART_INF_F;
  }
  for (int i = tid; i < W; i += kNmsThreads) { keptw[i] = 0u; remw[i] = 0u; }
  __syncthreads();

  // lower-triangle suppression bits: sup[i][w] bit l  <=>  iou(box_{32w+l}, box_i) > thresh, 32w+l < i.
  // Two rows per warp iteration share the loads of the column boxes.
  for (int ip = warp; 2 * ip < m; ip += kNmsThreads / 32) {
    const int i0 = 2 * ip, i1 = i0 + 1;
    const bool has1 = i1 < m;
    const float4 b0 = crn[i0], b1 = crn[has1 ? i1 : i0];
    const float q0 = qa[i0], q1 = qa[has1 ? i1 : i0];
    const int last = has1 ? i1 : i0;
    for (int w = 0; (w << 5) < last; ++w) {
      const int j = (w << 5) + lane;
      bool s0 = false, s1 = false, amb0 = false, amb1 = false;
      if (j < last) {
        const float4 bj = crn[j];
        const float qj = qa[j];
        {
          const float ex = fmaxf(0.f, fminf(b0.z, bj.z) - fmaxf(b0.x, bj.x));
          const float ey = fmaxf(0.f, fminf(b0.w, bj.w) - fmaxf(b0.y, bj.y));
          const float inter = ex * ey, r = q0 + qj;
          s0 = inter > r * 1.0001f;
          amb0 = !s0 && !(inter < r * 0.9999f);
        }
        {
          const float ex = fmaxf(0.f, fminf(b1.z, bj.z) - fmaxf(b1.x, bj.x));
          const float ey = fmaxf(0.f, fminf(b1.w, bj.w) - fmaxf(b1.y, bj.y));
          const float inter = ex * ey, r = q1 + qj;
          s1 = inter > r * 1.0001f;
          amb1 = !s1 && !(inter < r * 0.9999f);
        }
        if (!fast_ok) { amb0 = true; amb1 = true; }
        if (j >= i0) { s0 = false; amb0 = false; }
        if (!has1) { s1 = false; amb1 = false; }
      }
      if (__any_sync(SSDG_FULL, amb0 || amb1)) {
        // inside the margin (or no fast test): the formula itself, IEEE float32, no contraction
        if (amb0 || amb1) {
          const float4 bj = crn[j];
          const float aj = area[j];
          if (amb0) {
            const float ex = fmaxf(0.f, __fsub_rn(fminf(b0.z, bj.z), fmaxf(b0.x, bj.x)));
            const float ey = fmaxf(0.f, __fsub_rn(fminf(b0.w, bj.w), fmaxf(b0.y, bj.y)));
            const float inter = __fmul_rn(ex, ey);
            const float den = __fadd_rn(__fsub_rn(__fadd_rn(aj, area[i0]), inter), 1e-10f);
            s0 = __fdiv_rn(inter, den) > thr;
          }
          if (amb1) {
            const float ex = fmaxf(0.f, __fsub_rn(fminf(b1.z, bj.z), fmaxf(b1.x, bj.x)));
            const float ey = fmaxf(0.f, __fsub_rn(fminf(b1.w, bj.w), fmaxf(b1.y, bj.y)));
            const float inter = __fmul_rn(ex, ey);
            const float den = __fadd_rn(__fsub_rn(__fadd_rn(aj, area[i1]), inter), 1e-10f);
            s1 = __fdiv_rn(inter, den) > thr;
          }
        }
      }
      const u32 bits0 = __ballot_sync(SSDG_FULL, s0), bits1 = __ballot_sync(SSDG_FULL, s1);
      if (lane == 0) {
        if ((w << 5) < i0) sup[(size_t)i0 * W + w] = bits0;
        if (has1) sup[(size_t)i1 * W + w] = bits1;
      }
    }
  }
  __syncthreads();

  // fixed point of  kept(i) <=> no kept j < i with sup(i, j);  removed(i) <=> some kept j < i with sup(i, j)
  for (;;) {
    int unknown = 0;
    for (int i = tid; i < m; i += kNmsThreads) {
      const u32 bit = 1u << (i & 31);
      if ((keptw[i >> 5] | remw[i >> 5]) & bit) continue;
      bool hit_kept = false, all_removed = true;
      for (int w = 0; (w << 5) < i; ++w) {
        const u32 sb = sup[(size_t)i * W + w];
        if (sb & keptw[w]) hit_kept = true;
        if (sb & ~remw[w]) all_removed = false;
      }
      if (hit_kept) atomicOr(&remw[i >> 5], bit);
      else if (all_removed) atomicOr(&keptw[i >> 5], bit);
      else unknown = 1;
    }
    if (!__syncthreads_or(unknown)) break;
  }
  __syncthreads();

  // kept priors in visit order
  int* ok = P.out_kept + list * (size_t)P.top_k;
  float* os = P.out_score ? P.out_score + list * (size_t)P.top_k : nullptr;
  int total = 0;
  for (int w = 0; w < W; ++w) total += __popc(keptw[w]);
  for (int i = tid; i < P.top_k; i += kNmsThreads) {
    if (i >= total) { ok[i] = -1; if (os) os[i] = 0.f; }
  }
  for (int i = tid; i < m; i += kNmsThreads) {
    if (keptw[i >> 5] & (1u << (i & 31))) {
      int rank = __popc(keptw[i >> 5] & ((1u << (i & 31)) - 1u));
      for (int w = 0; w < (i >> 5); ++w) rank += __popc(keptw[w]);
      ok[rank] = (int)(~(u32)keys[i]);
      if (os) os[rank] = unkey32((u32)(keys[i] >> 32));
    }
  }
  if (tid == 0) P.out_count[list] = total;
}

static int f_warps_for(int C) {
  const size_t budget = 200 * 1024 - (size_t)kFWarps * 2 * C * 4;
  int w = (int)(budget / ((size_t)kFStages * 32 * C * 4));
  return w > kFWarps ? kFWarps : w;
}
static int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}
static size_t nms_smem_bytes(int sortn) {
  const int W = sortn / 32;
  return (size_t)sortn * (8 + 16 + 4 + 4) + (size_t)sortn * W * 4 + 2 * W * 4 + 256 * 4 + 64;
}

struct DetectWs {
  u32* ccount;
  u64* cand;
  float* boxes;
};
static size_t detect_ws_layout(long long batch, int A, int C, DetectWs* out, unsigned char* base, bool need_boxes) {
  size_t o = 0;
  const size_t lists = (size_t)batch * (C - 1);
  if (out) out->ccount = (u32*)(base + o);
  o += align_up(lists * 4, 256);
  if (out) out->cand = (u64*)(base + o);
  o += align_up(lists * (size_t)A * 8, 256);
  if (need_boxes) {
    if (out) out->boxes = (float*)(base + o);
    o += align_up((size_t)batch * A * 16, 256);
  }
  return o;
}

static int run_nms(const DetectWs& ws, const float* boxes, long long batch, int A, int C, int top_k, float iou_thresh,
                   int* out_kept, int* out_count, float* out_score, cudaStream_t st) {
  NmsParams Q;
  Q.ccount = ws.ccount; Q.cand = ws.cand; Q.boxes = boxes; Q.A = A; Q.n_fg = C - 1; Q.top_k = top_k;
  Q.sortn = next_pow2(top_k); Q.iou_thresh = iou_thresh;
  Q.out_kept = out_kept; Q.out_count = out_count; Q.out_score = out_score;
  const size_t smem = nms_smem_bytes(Q.sortn);
  if ((int)smem > max_smem_optin()) return SSDG_ERR_LIMIT;
  if (smem > 48 * 1024)
    SSDG_CUDA_TRY(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long lists = batch * (C - 1);
  if (lists > 0x7fffffffll) return SSDG_ERR_LIMIT;
  prof_begin(SSDG_PROF_NMS, st);
  nms_kernel<<<(unsigned)lists, kNmsThreads, smem, st>>>(Q);
  prof_end(SSDG_PROF_NMS, st);
  SSDG_LAUNCH_CHECK();
  return SSDG_OK;
}

}  // namespace ssdg

using namespace ssdg;

extern "C" size_t ssdg_detect_workspace_bytes(int64_t batch, int32_t n_priors, int32_t n_classes, int32_t top_k) {
  (void)top_k;
  if (batch <= 0 || n_priors <= 0 || n_classes < 2) return 0;
  return detect_ws_layout(batch, n_priors, n_classes, nullptr, nullptr, true);
}

extern "C" int ssdg_detect(const float* pred_cls, const float* pred_box, const void* priors, int32_t prior_dtype,
                           int64_t batch, int32_t n_priors, int32_t n_classes, float score_thresh, int32_t top_k,
                           float iou_thresh, int32_t* out_kept, int32_t* out_count, float* out_kept_score,
                           float* out_boxes, float* out_probs, float head_thresh, float* head_score,
                           int32_t* head_cls, uint8_t* head_mask, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (!pred_cls || !pred_box || !priors || !out_kept || !out_count) return SSDG_ERR_ARG;
  if (batch <= 0 || n_priors <= 0 || n_classes < 2 || top_k <= 0) return SSDG_ERR_ARG;
  if (prior_dtype != SSDG_F32 && prior_dtype != SSDG_F64) return SSDG_ERR_ARG;
  if (top_k > 1024) return SSDG_ERR_LIMIT;
  if (((uintptr_t)pred_cls | (uintptr_t)pred_box | (uintptr_t)priors | (uintptr_t)out_boxes | (uintptr_t)out_probs) & 15)
    return SSDG_ERR_ALIGN;
  if (!workspace || ((uintptr_t)workspace & 255) ||
      workspace_bytes < ssdg_detect_workspace_bytes(batch, n_priors, n_classes, top_k))
    return SSDG_ERR_WORKSPACE;
  const int warps = f_warps_for(n_classes);
  if (warps < 1) return SSDG_ERR_LIMIT;
  cudaStream_t st = (cudaStream_t)stream;
  DetectWs ws;
  detect_ws_layout(batch, n_priors, n_classes, &ws, (unsigned char*)workspace, true);
  DetectParams P;
  P.pred_cls = pred_cls; P.pred_box = pred_box; P.priors = priors;
  P.N = (long long)batch * n_priors; P.A = n_priors; P.C = n_classes; P.score_thresh = score_thresh;
  P.ccount = ws.ccount; P.cand = ws.cand; P.boxes = out_boxes ? out_boxes : ws.boxes; P.probs = out_probs;
  P.head_thresh = head_thresh; P.head_score = head_score; P.head_cls = head_cls; P.head_mask = head_mask;
  SSDG_CUDA_TRY(cudaMemsetAsync(ws.ccount, 0, (size_t)batch * (n_classes - 1) * 4, st));
  const size_t smem = (size_t)warps * kFStages * 32 * n_classes * 4 + kFWarps * kFStages * 8 +
                      (size_t)kFWarps * 2 * n_classes * 4 + 128;
  int grid = sm_count();
  const long long tiles = (P.N + 31) / 32;
  const long long need = (tiles + warps - 1) / warps;
  if (need < grid) grid = (int)need;
  prof_begin(SSDG_PROF_FILTER, st);
  if (prior_dtype == SSDG_F64) {
    SSDG_CUDA_TRY(cudaFuncSetAttribute(filter_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filter_kernel<double><<<grid, kFThreads, smem, st>>>(P, warps);
  } else {
    SSDG_CUDA_TRY(cudaFuncSetAttribute(filter_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filter_kernel<float><<<grid, kFThreads, smem, st>>>(P, warps);
  }
  prof_end(SSDG_PROF_FILTER, st);
  SSDG_LAUNCH_CHECK();
  return run_nms(ws, P.boxes, batch, n_priors, n_classes, top_k, iou_thresh, out_kept, out_count, out_kept_score, st);
}

extern "C" int ssdg_nms(const float* probs, const float* boxes, int64_t batch, int32_t n_priors, int32_t n_classes,
                        float score_thresh, int32_t top_k, float iou_thresh, int32_t* out_kept, int32_t* out_count,
                        float* out_kept_score, void* workspace, size_t workspace_bytes, void* stream) {
  if (!probs || !boxes || !out_kept || !out_count) return SSDG_ERR_ARG;
  if (batch <= 0 || n_priors <= 0 || n_classes < 2 || top_k <= 0) return SSDG_ERR_ARG;
  if (top_k > 1024) return SSDG_ERR_LIMIT;
  if ((uintptr_t)boxes & 15) return SSDG_ERR_ALIGN;
  if (!workspace || ((uintptr_t)workspace & 255) ||
      workspace_bytes < ssdg_detect_workspace_bytes(batch, n_priors, n_classes, top_k))
    return SSDG_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  DetectWs ws;
  detect_ws_layout(batch, n_priors, n_classes, &ws, (unsigned char*)workspace, true);
  DetectParams P;
  P.pred_cls = probs; P.pred_box = nullptr; P.priors = nullptr;
  P.N = (long long)batch * n_priors; P.A = n_priors; P.C = n_classes; P.score_thresh = score_thresh;
  P.ccount = ws.ccount; P.cand = ws.cand; P.boxes = nullptr; P.probs = nullptr;
  P.head_thresh = 0.f; P.head_score = nullptr; P.head_cls = nullptr; P.head_mask = nullptr;
  SSDG_CUDA_TRY(cudaMemsetAsync(ws.ccount, 0, (size_t)batch * (n_classes - 1) * 4, st));
  int grid = sm_count() * 8;
  emit_kernel<<<grid, 256, 0, st>>>(P);
  SSDG_LAUNCH_CHECK();
  return run_nms(ws, boxes, batch, n_priors, n_classes, top_k, iou_thresh, out_kept, out_count, out_kept_score, st);
}
