// Post-processing: softmax score head (models/ssd_model.py:479-488), box decode (:466-467) and the
// per-class score-threshold / top-k / greedy NMS the reference lacks (spec: oracle/ssd_oracle.py
// nms_per_class, IoU formula utils/bbox.py:13-25 in float32).
//
//   filter_kernel  one streaming pass over the logits [B,A,C].  Tiles are 32 priors of ONE image
//                  (image-aligned), each warp owns a ring of tiles filled by 1-D bulk TMA
//                  (cp.async.bulk + mbarrier); lane r owns row r in shared memory (stride C words:
//                  conflict-free for odd C).  Candidates (p > score_thresh) are found with one ballot
//                  per class; a warp prefix scan over the classes gives every class its slot range
//                  in the tile's private segment of the candidate buffer, so there is NO atomic and
//                  no contention: the tile writes  cand[tile][class-major]  and one packed
//                  (offset,count) word per class into  meta[image][class][tile].
//   nms_kernel     one CTA per (image, class): coalesced read + block scan of the class's meta row,
//                  gather of the candidates, exact top-k by (score desc, prior asc) -- radix select when
//                  the list is longer than the sort width, then a bitonic sort whose short strides are
//                  warp-local -- the lower-triangle suppression bit matrix built by 32x32 tasks in
//                  registers (division-free margin test, the exact IEEE formula only inside the margin),
//                  and a parallel fixed-point resolution of "kept(i) <=> no kept j < i suppresses i".
#include <math_constants.h>
#include "common.cuh"

namespace ssdg {

constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFStages = 2;
constexpr int kNmsThreads = 256;
constexpr int kNmsWarps = kNmsThreads / 32;

struct DetectParams {
  const float* pred_cls;   // logits (filter) or probabilities (kProbs)
  const float* pred_box;
  const void* priors;
  int B, A, C, tpi;        // tpi: tiles per image
  int tma_ok;              // every tile start / size is 16-byte aligned
  float score_thresh;
  u32* meta;               // [B][C-1][tpi]  (offset << 8) | count
  u64* cand;               // [B*tpi][32*(C-1)]
  float* boxes;            // [B*A,4] decoded
  float* probs;            // optional [B*A,C]
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

// One warp tile: `rows` priors of image b starting at prior 32*j; lane r owns row r (rows may be
// overwritten with the probabilities).  kProbs: the rows already hold probabilities.
template <typename TP, bool kProbs>
__device__ __forceinline__ void filter_tile(const DetectParams& P, int b, int j, int rows, float* tile, u32* wmask,
                                            u32* wbase, int lane) {
  const int C = P.C, nfg = P.C - 1;
  const bool valid = lane < rows;
  const int a = j * 32 + lane;
  const long long n = (long long)b * P.A + a;
  float* row = tile + (size_t)(valid ? lane : 0) * C;
  float m = 0.f, s = 1.f;
  if (!kProbs && valid) {
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
  // p_c > thresh  <=>  x_c - m > log(thresh * s): pre-filter in logit space with a margin, then the
  // exact score  exp(x_c - m) / s  decides.
  const float cut = (P.score_thresh > 0.f) ? __logf(P.score_thresh * s) - 1e-3f : -CUDART_INF_F;
  for (int c = 0; c < nfg; ++c) {
    bool p = false;
    if (valid) {
      if (kProbs) {
        p = row[c] > P.score_thresh;
      } else {
        const float d = row[c] - m;
        if (d > cut) p = __fdiv_rn(__expf(d), s) > P.score_thresh;
      }
    }
    const u32 mc = __ballot_sync(SSDG_FULL, p);
    if (lane == 0) wmask[c] = mc;
  }
  __syncwarp();
  // slot ranges: exclusive prefix sum of the per-class counts, classes in ascending order
  {
    u32* mrow = P.meta + (size_t)b * nfg * P.tpi + j;
    int running = 0;
    for (int c0 = 0; c0 < nfg; c0 += 32) {
      const int c = c0 + lane;
      const int cnt = c < nfg ? __popc(wmask[c]) : 0;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(SSDG_FULL, incl, o);
        if (lane >= o) incl += v;
      }
      const int off = running + incl - cnt;
      if (c < nfg) {
        wbase[c] = (u32)off;
        mrow[(size_t)c * P.tpi] = ((u32)off << 8) | (u32)cnt;
      }
      running += __shfl_sync(SSDG_FULL, incl, 31);
    }
  }
  __syncwarp();
  {
    u64* seg = P.cand + ((size_t)b * P.tpi + j) * (size_t)(32 * nfg);
    const u32 lt = (1u << lane) - 1u;
    for (int c = 0; c < nfg; ++c) {
      const u32 mc = wmask[c];
      if (!mc) continue;
      if ((mc >> lane) & 1u) {
        const float score = kProbs ? row[c] : __fdiv_rn(__expf(row[c] - m), s);
        seg[wbase[c] + (u32)__popc(mc & lt)] = ((u64)key32(score) << 32) | (u64)(~(u32)a);
      }
    }
  }
  if (kProbs || !valid) return;
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

template <typename TP, bool kProbs>
__global__ void __launch_bounds__(kFThreads, 1) filter_kernel(DetectParams P, int warps_per_cta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = P.C, A = P.A, tpi = P.tpi;
  const size_t tile_floats = (size_t)32 * C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* bufs = reinterpret_cast<float*>(smem_raw);
  u64* bars = reinterpret_cast<u64*>(smem_raw + (size_t)warps_per_cta * kFStages * tile_floats * 4);
  u32* scratch = reinterpret_cast<u32*>(bars + kFWarps * kFStages);   // [warps][2][C]
  if (tid == 0) {
    for (int i = 0; i < warps_per_cta * kFStages; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (warp >= warps_per_cta) return;
  u32* wmask = scratch + (size_t)warp * 2 * C;
  u32* wbase = wmask + C;
  const long long ntiles = (long long)P.B * tpi;
  const long long gw = (long long)blockIdx.x * warps_per_cta + warp;
  const long long stride = (long long)gridDim.x * warps_per_cta;
  float* mybuf = bufs + (size_t)warp * kFStages * tile_floats;
  u64* mybar = bars + warp * kFStages;
  const u64 pol = evict_first_policy();

  auto tile_rows = [&](long long t) { const int j = (int)(t % tpi); return min(32, A - j * 32); };
  auto tile_src = [&](long long t) {
    const long long b = t / tpi;
    const int j = (int)(t - b * tpi);
    return P.pred_cls + ((size_t)b * A + (size_t)j * 32) * C;
  };
  auto issue = [&](long long t, int s) {  // lane 0 only
    const u32 bytes = (u32)tile_rows(t) * (u32)C * 4u;
    mbar_arrive_expect_tx(&mybar[s], bytes);
    tma_load_hint(mybuf + (size_t)s * tile_floats, tile_src(t), bytes, &mybar[s], pol);
  };

  if (P.tma_ok && lane == 0)
    for (int s = 0; s < kFStages; ++s) {
      const long long t = gw + (long long)s * stride;
      if (t < ntiles) issue(t, s);
    }
  int k = 0;
  for (long long t = gw; t < ntiles; t += stride, ++k) {
    const int s = k % kFStages;
    float* tile = mybuf + (size_t)s * tile_floats;
    const int rows = tile_rows(t);
    const int b = (int)(t / tpi), j = (int)(t - (long long)b * tpi);
    if (P.tma_ok) {
      mbar_wait(&mybar[s], (u32)((k / kFStages) & 1));
    } else {  // unaligned shapes: plain cooperative copy
      const float* g = tile_src(t);
      for (int i = lane; i < rows * C; i += 32) tile[i] = g[i];
      __syncwarp();
    }
    filter_tile<TP, kProbs>(P, b, j, rows, tile, wmask, wbase, lane);
    __syncwarp();
    if (!kProbs && P.probs) {  // the tile layout in shared memory equals the layout in global memory
      float* dst = P.probs + ((size_t)b * A + (size_t)j * 32) * C;
      for (int i = lane; i < rows * C; i += 32) __stcs(&dst[i], tile[i]);
      __syncwarp();
    }
    const long long tn = t + (long long)kFStages * stride;
    if (P.tma_ok && lane == 0 && tn < ntiles) issue(tn, s);
  }
}

// ---- per-(image, class) NMS ---------------------------------------------------------------------------
struct NmsParams {
  const u32* meta;
  const u64* cand;
  const float* boxes;  // [B,A,4]
  int A, n_fg, tpi, top_k, sortn;  // sortn: power of two >= max(top_k, 32)
  float iou_thresh;
  int* out_kept;
  int* out_count;
  float* out_score;
};

// Bitonic sort (descending) of n = 2^k keys in shared memory.  Pair i of a stage is handled by thread
// i (mod blockDim); for strides <= 32 all pairs of a warp live in that warp's own 64-key block, so
// only the long strides need a block-wide barrier.
__device__ __forceinline__ void bitonic_sort_desc(u64* keys, int n, int tid) {
  __syncthreads();
  bool wide_prev = true;
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const bool wide = stride > 32;
      if (wide || wide_prev) __syncthreads(); else __syncwarp();
      wide_prev = wide;
      for (int i = tid; i < (n >> 1); i += kNmsThreads) {
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
  const int sortn = P.sortn, W = sortn >> 5, WP = W | 1;
  u64* keys = reinterpret_cast<u64*>(smem_raw);                 // [sortn]
  float4* crn = reinterpret_cast<float4*>(keys + sortn);        // [sortn] x1,y1,x2,y2
  float2* q2 = reinterpret_cast<float2*>(crn + sortn);          // [sortn] qa*1.0001, qa*0.9999
  float* area = reinterpret_cast<float*>(q2 + sortn);           // [sortn]
  u32* sup = reinterpret_cast<u32*>(area + sortn);              // [sortn][WP] lower triangle
  u32* keptw = sup + (size_t)sortn * WP;                        // [W]
  u32* remw = keptw + W;                                        // [W]
  u32* hist = remw + W;                                         // [256]
  int* wsum = reinterpret_cast<int*>(hist + 256);               // [kNmsWarps]
  __shared__ u64 sel_prefix;
  __shared__ int sel_k, sel_fill;

  const size_t list = blockIdx.x;
  const int b = (int)(list / P.n_fg);
  const int tpi = P.tpi;
  const u32* mrow = P.meta + list * (size_t)tpi;
  const u64* cbase = P.cand + (size_t)b * tpi * (size_t)(32 * P.n_fg);

  // candidates of this class: per-tile (offset, count) -> exclusive scan -> total
  int mine = 0;
  for (int j = tid; j < tpi; j += kNmsThreads) mine += (int)(mrow[j] & 255u);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(SSDG_FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  for (int i = tid; i < sortn; i += kNmsThreads) keys[i] = 0ull;
  __syncthreads();
  int before = incl - mine, n = 0;
  for (int w = 0; w < kNmsWarps; ++w) {
    if (w < warp) before += wsum[w];
    n += wsum[w];
  }

  if (n <= sortn) {
    int dst = before;
    for (int j = tid; j < tpi; j += kNmsThreads) {
      const u32 mv = mrow[j];
      const u64* src = cbase + (size_t)j * (32 * P.n_fg) + (mv >> 8);
      for (int e = 0; e < (int)(mv & 255u); ++e) keys[dst++] = src[e];
    }
  } else {
    // Radix select of the top_k-th largest composite key (keys are unique), 8 bits per pass.
    if (tid == 0) { sel_prefix = 0ull; sel_k = P.top_k; sel_fill = 0; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += kNmsThreads) hist[i] = 0u;
      __syncthreads();
      const u64 pre = sel_prefix;
      const u64 hmask = shift == 56 ? 0ull : (~0ull << (shift + 8));
      for (int j = tid; j < tpi; j += kNmsThreads) {
        const u32 mv = mrow[j];
        const u64* src = cbase + (size_t)j * (32 * P.n_fg) + (mv >> 8);
        for (int e = 0; e < (int)(mv & 255u); ++e) {
          const u64 v = src[e];
          if ((v & hmask) == pre) atomicAdd(&hist[(u32)(v >> shift) & 255u], 1u);
        }
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
    for (int j = tid; j < tpi; j += kNmsThreads) {
      const u32 mv = mrow[j];
      const u64* src = cbase + (size_t)j * (32 * P.n_fg) + (mv >> 8);
      for (int e = 0; e < (int)(mv & 255u); ++e) {
        const u64 v = src[e];
        if (v >= kth) {
          const int p = atomicAdd(&sel_fill, 1);
          if (p < sortn) keys[p] = v;
        }
      }
    }
    __syncthreads();
    n = min(sel_fill, sortn);
  }
  int sn = 32;
  while (sn < n) sn <<= 1;           // n <= sortn here; the keys beyond n are 0 and sort to the end
  bitonic_sort_desc(keys, sn, tid);
  const int m = min(n, P.top_k);
  const int mpad = (m + 31) & ~31;

  // Decoded boxes; corners in the formula's own float32 operations (utils/bbox.py:13-21).
  // qa = q*(area + 0.5e-10), q = thr/(1+thr):  iou > thr  <=>  inter > qa_i + qa_j  in exact arithmetic
  // (positive denominator); boxes that can never overlap anything (w <= 0, h <= 0, non-finite, padding)
  // get qa = +inf so the fast test rejects them exactly like the formula does (their intersection is 0).
  const float thr = P.iou_thresh;
  const bool fast_ok = thr > 0.f && thr < 1e6f;
  const float q = fast_ok ? thr / (1.f + thr) : 0.f;
  for (int i = tid; i < mpad; i += kNmsThreads) {
    float4 cr = make_float4(0.f, 0.f, 0.f, 0.f);
    float ar = 0.f, qa = CUDART_INF_F;
    if (i < m) {
      const int a = (int)(~(u32)keys[i]);
      const float4 bx = __ldg(reinterpret_cast<const float4*>(P.boxes) + (size_t)b * P.A + a);
      const float hw = __fmul_rn(bx.z, 0.5f), hh = __fmul_rn(bx.w, 0.5f);
      cr = make_float4(__fsub_rn(bx.x, hw), __fsub_rn(bx.y, hh), __fadd_rn(bx.x, hw), __fadd_rn(bx.y, hh));
      ar = __fmul_rn(bx.z, bx.w);
      const bool sane = bx.z > 0.f && bx.w > 0.f && isfinite(cr.x) && isfinite(cr.y) && isfinite(cr.z) &&
                        isfinite(cr.w) && isfinite(ar);
      if (sane) qa = q * (ar + 0.5e-10f);
    }
    crn[i] = cr;
    area[i] = ar;
    q2[i] = make_float2(qa * 1.0001f, qa * 0.9999f);
  }
  for (int i = tid; i < W; i += kNmsThreads) { keptw[i] = 0u; remw[i] = 0u; }
  __syncthreads();

  // Suppression bits, lower triangle: sup[i][w] bit l  <=>  iou(box_{32w+l}, box_i) > thresh, 32w+l < i.
  // Task (g, w<=g): lane = row 32g+lane, loop over the 32 columns of group w (uniform shared loads).
  const int ngroups = mpad >> 5;
  const int ntasks = ngroups * (ngroups + 1) / 2;
  for (int task = warp; task < ntasks; task += kNmsWarps) {
    int g = 0;
    while ((g + 1) * (g + 2) / 2 <= task) ++g;
    const int w = task - g * (g + 1) / 2;
    const int i = (g << 5) + lane;
    const float4 bi = crn[i];
    const float2 qi = q2[i];
    u32 bits = 0u, amb = 0u;
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const int j = (w << 5) + jj;
      const float4 bj = crn[j];
      const float2 qj = q2[j];
      const float ex = fminf(bi.z, bj.z) - fmaxf(bi.x, bj.x);
      const float ey = fminf(bi.w, bj.w) - fmaxf(bi.y, bj.y);
      const float inter = ex * ey;
      const bool over = ex > 0.f && ey > 0.f;
      const bool s = over && inter > qi.x + qj.x;
      const bool am = over && !s && !(inter < qi.y + qj.y);
      bits |= (s ? 1u : 0u) << jj;
      amb |= (am ? 1u : 0u) << jj;
    }
    if (!fast_ok) amb = 0xffffffffu;
    if (w == g) { const u32 lower = (1u << lane) - 1u; bits &= lower; amb &= lower; }
    if (i >= m) { bits = 0u; amb = 0u; }
    // inside the margin (or no fast test): the formula itself, IEEE float32, no contraction
    while (amb) {
      const int jj = __ffs(amb) - 1;
      amb &= amb - 1;
      const int j = (w << 5) + jj;
      if (j >= m) continue;
      const float4 bj = crn[j];
      const float ex = fmaxf(0.f, __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
      const float ey = fmaxf(0.f, __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
      const float inter = __fmul_rn(ex, ey);
      const float den = __fadd_rn(__fsub_rn(__fadd_rn(area[j], area[i]), inter), 1e-10f);
      if (__fdiv_rn(inter, den) > thr) bits |= 1u << jj; else bits &= ~(1u << jj);
    }
    sup[(size_t)i * WP + w] = bits;
  }
  __syncthreads();

  // fixed point of  kept(i) <=> no kept j < i with sup(i, j);  removed(i) <=> some kept j < i with sup(i, j)
  for (;;) {
    int unknown = 0;
    for (int i = tid; i < m; i += kNmsThreads) {
      const u32 bit = 1u << (i & 31);
      if ((keptw[i >> 5] | remw[i >> 5]) & bit) continue;
      bool hit_kept = false, all_removed = true;
      for (int w = 0; w <= (i >> 5); ++w) {
        const u32 sb = sup[(size_t)i * WP + w];
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
  const int W = sortn / 32, WP = W | 1;
  return (size_t)sortn * (8 + 16 + 8 + 4) + (size_t)sortn * WP * 4 + 2 * W * 4 + 256 * 4 + kNmsWarps * 4 + 128;
}

struct DetectWs {
  u32* meta;
  u64* cand;
  float* boxes;
};
static size_t detect_ws_layout(long long batch, int A, int C, DetectWs* out, unsigned char* base) {
  size_t o = 0;
  const size_t tpi = ((size_t)A + 31) / 32;
  const size_t nfg = (size_t)C - 1;
  if (out) out->meta = (u32*)(base + o);
  o += align_up((size_t)batch * nfg * tpi * 4, 256);
  if (out) out->cand = (u64*)(base + o);
  o += align_up((size_t)batch * tpi * 32 * nfg * 8, 256);
  if (out) out->boxes = (float*)(base + o);
  o += align_up((size_t)batch * A * 16, 256);
  return o;
}

template <bool kProbs>
static int run_filter(DetectParams& P, int prior_dtype, cudaStream_t st) {
  const int warps = f_warps_for(P.C);
  if (warps < 1) return SSDG_ERR_LIMIT;
  const size_t smem = (size_t)warps * kFStages * 32 * P.C * 4 + kFWarps * kFStages * 8 + (size_t)kFWarps * 2 * P.C * 4 + 128;
  int grid = sm_count();
  const long long tiles = (long long)P.B * P.tpi;
  const long long need = (tiles + warps - 1) / warps;
  if (need < grid) grid = (int)need;
  P.tma_ok = (((long long)P.A * P.C) % 4 == 0) && (((uintptr_t)P.pred_cls & 15) == 0);
  prof_begin(SSDG_PROF_FILTER, st);
  if (prior_dtype == SSDG_F64) {
    SSDG_CUDA_TRY(cudaFuncSetAttribute(filter_kernel<double, kProbs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filter_kernel<double, kProbs><<<grid, kFThreads, smem, st>>>(P, warps);
  } else {
    SSDG_CUDA_TRY(cudaFuncSetAttribute(filter_kernel<float, kProbs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filter_kernel<float, kProbs><<<grid, kFThreads, smem, st>>>(P, warps);
  }
  prof_end(SSDG_PROF_FILTER, st);
  SSDG_LAUNCH_CHECK();
  return SSDG_OK;
}

static int run_nms(const DetectWs& ws, const float* boxes, long long batch, int A, int C, int top_k, float iou_thresh,
                   int* out_kept, int* out_count, float* out_score, cudaStream_t st) {
  NmsParams Q;
  Q.meta = ws.meta; Q.cand = ws.cand; Q.boxes = boxes; Q.A = A; Q.n_fg = C - 1; Q.tpi = (A + 31) / 32; Q.top_k = top_k;
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
  return detect_ws_layout(batch, n_priors, n_classes, nullptr, nullptr);
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
  if (top_k > 1024 || batch > 0x7fffffff / 64) return SSDG_ERR_LIMIT;
  if (((uintptr_t)pred_box | (uintptr_t)priors | (uintptr_t)out_boxes) & 15) return SSDG_ERR_ALIGN;
  if (((uintptr_t)pred_cls | (uintptr_t)out_probs) & 3) return SSDG_ERR_ALIGN;
  if (!workspace || ((uintptr_t)workspace & 255) ||
      workspace_bytes < ssdg_detect_workspace_bytes(batch, n_priors, n_classes, top_k))
    return SSDG_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  DetectWs ws;
  detect_ws_layout(batch, n_priors, n_classes, &ws, (unsigned char*)workspace);
  DetectParams P;
  P.pred_cls = pred_cls; P.pred_box = pred_box; P.priors = priors;
  P.B = (int)batch; P.A = n_priors; P.C = n_classes; P.tpi = (n_priors + 31) / 32; P.score_thresh = score_thresh;
  P.meta = ws.meta; P.cand = ws.cand; P.boxes = out_boxes ? out_boxes : ws.boxes; P.probs = out_probs;
  P.head_thresh = head_thresh; P.head_score = head_score; P.head_cls = head_cls; P.head_mask = head_mask;
  int rc = run_filter<false>(P, prior_dtype, st);
  if (rc) return rc;
  return run_nms(ws, P.boxes, batch, n_priors, n_classes, top_k, iou_thresh, out_kept, out_count, out_kept_score, st);
}

extern "C" int ssdg_nms(const float* probs, const float* boxes, int64_t batch, int32_t n_priors, int32_t n_classes,
                        float score_thresh, int32_t top_k, float iou_thresh, int32_t* out_kept, int32_t* out_count,
                        float* out_kept_score, void* workspace, size_t workspace_bytes, void* stream) {
  if (!probs || !boxes || !out_kept || !out_count) return SSDG_ERR_ARG;
  if (batch <= 0 || n_priors <= 0 || n_classes < 2 || top_k <= 0) return SSDG_ERR_ARG;
  if (top_k > 1024 || batch > 0x7fffffff / 64) return SSDG_ERR_LIMIT;
  if ((uintptr_t)boxes & 15) return SSDG_ERR_ALIGN;
  if ((uintptr_t)probs & 3) return SSDG_ERR_ALIGN;
  if (!workspace || ((uintptr_t)workspace & 255) ||
      workspace_bytes < ssdg_detect_workspace_bytes(batch, n_priors, n_classes, top_k))
    return SSDG_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  DetectWs ws;
  detect_ws_layout(batch, n_priors, n_classes, &ws, (unsigned char*)workspace);
  DetectParams P;
  P.pred_cls = probs; P.pred_box = nullptr; P.priors = nullptr;
  P.B = (int)batch; P.A = n_priors; P.C = n_classes; P.tpi = (n_priors + 31) / 32; P.score_thresh = score_thresh;
  P.meta = ws.meta; P.cand = ws.cand; P.boxes = nullptr; P.probs = nullptr;
  P.head_thresh = 0.f; P.head_score = nullptr; P.head_cls = nullptr; P.head_mask = nullptr;
  int rc = run_filter<true>(P, SSDG_F32, st);
  if (rc) return rc;
  return run_nms(ws, boxes, batch, n_priors, n_classes, top_k, iou_thresh, out_kept, out_count, out_kept_score, st);
}
