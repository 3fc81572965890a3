// sb_moments.cu — the Gram matrix ΘᵀΘ of a polynomial library as a MOMENT matrix.
//
// Reference: `solve_SINDy_one_step` stacks Θ(x) (N×K) and runs LAPACK on it (`sindy.py:260-288`); its normal
// equations only need G = ΘᵀΘ. For a polynomial library G[k,l] = Σ_n x_n^(α_k+α_l): every entry is one of the
// C(d+2p, d) power sums of degree ≤ 2p (286 for d = 3, p = 5 against K(K+1)/2 = 1596 distinct dense entries), and
// the same moments give the linear Lie-derivative regulariser Σ_v tr(A_v G A_vᵀ) (`train.py:503-507`) for any
// number of generators. Tensor cores would need a 3×TF32 split of a 64×64 syrk whose useful work is ~5× larger
// than these power sums; the moments run on the FP32 FMA pipe at one packed FMA per two moments.
//
// Per sample: leading powers x0^a (a ≤ 2p) and the table T of the monomials of the trailing d−1 variables
// (degree ≤ 2p, graded order, so "degree ≤ 2p−a" is a PREFIX of T); moment(a, j) += x0^a · T[j] as
// fma.rn.f32x2 over adjacent j. The 286 accumulators of (3,5) do not fit one thread: the CTA is warp-specialised,
// warps 0-3 own a ∈ [0,2), warps 4-7 a ∈ [2,10] (one warp of each per scheduler), both sweep the same TMA-staged x tile (x is read once, 4·d
// bytes per sample). fp32 per thread, fp64 across threads, ordered last-block reduction (deterministic). A tiny
// gather kernel then writes the K×K Gram from the moments.
#include <atomic>
#include <mutex>

#include "sb_common.cuh"
#include "sb_tma.cuh"

namespace sb {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 2048;  // samples per stage (x only); 4096 x 2 measured slower (K = 56: 1.70 against 1.59 ms)
constexpr int kStages = 4;

constexpr int trailing_len(int dt, int deg) { return deg < 0 ? 0 : n_poly_terms(dt, deg); }

// One role: leading exponents a in [A_LO, A_HI] of a library with D variables and degree P (moments to 2P).
template <int D, int P, int A_LO, int A_HI>
struct Role {
  static constexpr int Q = 2 * P;
  static constexpr int DT = D - 1;
  static constexpr int NTR = trailing_len(DT, Q - A_LO);  // trailing monomials this role needs
  static constexpr int len(int a) { return trailing_len(DT, Q - a); }
  static constexpr int pair_off(int a) {  // first accumulator pair of exponent a
    int o = 0;
    for (int q = A_LO; q < a; ++q) o += (len(q) + 1) / 2;
    return o;
  }
  static constexpr int NPAIR = pair_off(A_HI + 1);
  static constexpr int out_off(int a) {  // first moment index of exponent a within this role's chunk
    int o = 0;
    for (int q = A_LO; q < a; ++q) o += len(q);
    return o;
  }
  static constexpr int NOUT = out_off(A_HI + 1);

  __device__ static __forceinline__ void zero(float2 (&acc)[NPAIR]) {
    static_for<0, NPAIR>([&](auto i) { acc[i] = make_float2(0.f, 0.f); });
  }

  __device__ static __forceinline__ void accumulate(const float (&x)[D], float2 (&acc)[NPAIR]) {
    float t[NTR];
    if constexpr (DT >= 1) {
      float xt[DT > 0 ? DT : 1];
      static_for<0, DT>([&](auto j) { xt[j] = x[1 + j]; });
      expand_poly<DT, Q - A_LO>(xt, t);
    } else {
      t[0] = 1.f;
    }
    float xa = 1.f;
    static_for<0, A_LO>([&](auto) { xa *= x[0]; });
    static_for<A_LO, A_HI + 1>([&](auto ac) {
      constexpr int a = ac;
      constexpr int L = len(a);
      constexpr int po = pair_off(a);
      const float2 xa2 = make_float2(xa, xa);
      static_for<0, (L + 1) / 2>([&](auto jc) {
        constexpr int j = jc;
        // the odd lane of the last pair may hold a monomial beyond this exponent's prefix: its sum is never read
        const float2 t2 = make_float2(t[2 * j], (2 * j + 1 < NTR) ? t[2 * j + 1] : 0.f);
        acc[po + j] = __ffma2_rn(xa2, t2, acc[po + j]);
      });
      if constexpr (a < A_HI) xa *= x[0];
    });
  }

  // warp-reduce every moment of this role into red[0 .. NOUT)
  __device__ static __forceinline__ void reduce(const float2 (&acc)[NPAIR], float* red, int lane) {
    static_for<A_LO, A_HI + 1>([&](auto ac) {
      constexpr int a = ac;
      constexpr int L = len(a);
      static_for<0, L>([&](auto jc) {
        constexpr int j = jc;
        const float v = warp_sum((j % 2 == 0) ? acc[pair_off(a) + j / 2].x : acc[pair_off(a) + j / 2].y);
        if (lane == 0) red[out_off(a) + j] = v;
      });
    });
  }
};

template <int D, int P>
struct MomCfg {
  static constexpr int Q = 2 * P;
  static constexpr int NM = n_poly_terms(D, Q);                    // all moments
  static constexpr bool kSplit = NM > 160;                         // two roles when one thread cannot hold them
  static constexpr int kSplitA = 2;                                // role 0: a < kSplitA, role 1: a >= kSplitA
  using R0 = Role<D, P, 0, kSplit ? kSplitA - 1 : Q>;
  using R1 = Role<D, P, kSplit ? kSplitA : Q, Q>;
  static constexpr int kMaxOut = kSplit ? (R0::NOUT > R1::NOUT ? R0::NOUT : R1::NOUT) : R0::NOUT;
  static constexpr size_t kSmemData = (size_t)kStages * kTile * D * sizeof(float);
  static constexpr size_t kSmemBytes = kSmemData + 2 * kStages * sizeof(uint64_t) + 16;
};

struct MomArgs {
  const float* x;
  int64_t n;
  int64_t n_bulk;
  int64_t n_tiles;
  double* partial;  // [grid][NM]
  unsigned int* ticket;
  double* moments;  // [NM], ordered by leading exponent a, then trailing graded index
};

// the tile loop of one role; `sub`/`nsub` = this warp's index / count among the warps of the role
template <class R, int D>
__device__ __forceinline__ void sweep(const MomArgs& a, const float* tiles, uint64_t* full, uint64_t* empty, int sub,
                                      int nsub, int lane, bool producer, float2 (&acc)[R::NPAIR]) {
  constexpr int kTileFloats = kTile * D;
  // one contiguous sample range per CTA, cut in quads (16-byte aligned TMA sources), equal to within 4 samples; the
  // first tile travels alone and the rest of the ring follows when it has landed (see sb_fused.cu)
  const int64_t quads = a.n_bulk >> 2;
  const int64_t s_begin = 4 * (quads * (int64_t)blockIdx.x / (int64_t)gridDim.x);
  const int64_t s_end = 4 * (quads * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x);
  const int my_tiles = (int)((s_end - s_begin + kTile - 1) / kTile);
  auto tile_count = [&](int t) -> int {
    const int64_t rem = s_end - s_begin - (int64_t)t * kTile;
    return (int)(rem < kTile ? rem : kTile);
  };
  auto issue = [&](int t, int stage) {
    const uint32_t bytes = (uint32_t)tile_count(t) * D * sizeof(float);
    mbar_expect_tx(&full[stage], bytes);
    tma_load_1d(const_cast<float*>(tiles) + (size_t)stage * kTileFloats, a.x + (s_begin + (int64_t)t * kTile) * D, bytes,
                &full[stage]);
  };
  if (producer && my_tiles > 0) issue(0, 0);
  for (int it = 0; it < my_tiles; ++it) {
    const int stage = it % kStages;
    if (producer && it > 0) {
      const int next = it + kStages - 1;
      if (next < my_tiles) {
        const int ps = (it - 1) % kStages;
        mbar_wait(&empty[ps], (uint32_t)((it - 1) / kStages) & 1u);
        issue(next, ps);
      }
    }
    mbar_wait(&full[stage], (uint32_t)(it / kStages) & 1u);
    if (producer && it == 0) {
      for (int s = 1; s < kStages; ++s)
        if (s < my_tiles) issue(s, s);
    }
    const float* sx = tiles + (size_t)stage * kTileFloats;
    const int cnt = tile_count(it);
#pragma unroll 1
    for (int j = sub * 32 + lane; j < cnt; j += nsub * 32) {
      float xs[D];
      static_for<0, D>([&](auto q) { xs[q] = sx[j * D + q]; });
      R::accumulate(xs, acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
  // ragged tail (n % 4 samples): first lanes of the role's first warp of block 0
  if (blockIdx.x == 0 && sub == 0) {
    const int64_t j = a.n_bulk + lane;
    if (j < a.n) {
      float xs[D];
      static_for<0, D>([&](auto q) { xs[q] = __ldg(a.x + j * D + q); });
      R::accumulate(xs, acc);
    }
  }
}

template <int D, int P>
__global__ void __launch_bounds__(kThreads, 1) moments_kernel(MomArgs a) {
  using C = MomCfg<D, P>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + C::kSmemData);
  uint64_t* empty = full + kStages;
  __shared__ float red[kWarps][C::kMaxOut];
  __shared__ int is_last;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
    fence_mbar_init();
  }
  __syncthreads();

  if constexpr (C::kSplit) {
    // warps 0-3 take role 0, warps 4-7 role 1: every scheduler (warp id mod 4) hosts one warp of each role, so the
    // unequal per-sample costs of the two roles add up identically on the four FMA pipes
    constexpr int kHalf = kWarps / 2;
    if (wid < kHalf) {
      float2 acc[C::R0::NPAIR];
      C::R0::zero(acc);
      sweep<typename C::R0, D>(a, tiles, full, empty, wid, kHalf, lane, tid == 0, acc);
      C::R0::reduce(acc, red[wid], lane);
    } else {
      float2 acc[C::R1::NPAIR];
      C::R1::zero(acc);
      sweep<typename C::R1, D>(a, tiles, full, empty, wid - kHalf, kHalf, lane, false, acc);
      C::R1::reduce(acc, red[wid], lane);
    }
  } else {
    float2 acc[C::R0::NPAIR];
    C::R0::zero(acc);
    sweep<typename C::R0, D>(a, tiles, full, empty, wid, kWarps, lane, tid == 0, acc);
    C::R0::reduce(acc, red[wid], lane);
  }
  __syncthreads();

  double* mine = a.partial + (int64_t)blockIdx.x * C::NM;
  for (int e = tid; e < C::NM; e += kThreads) {
    double v = 0.0;
    if constexpr (C::kSplit) {
      const int role = (e < C::R0::NOUT) ? 0 : 1;
      const int local = role ? e - C::R0::NOUT : e;
#pragma unroll
      for (int wq = 0; wq < kWarps / 2; ++wq) v += (double)red[role * (kWarps / 2) + wq][local];
    } else {
#pragma unroll
      for (int wq = 0; wq < kWarps; ++wq) v += (double)red[wq][e];
    }
    mine[e] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1u);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ordered, with 8 loads in flight per thread (the partials sit in L2: latency-bound) and a fixed association
  for (int e = tid; e < C::NM; e += kThreads) {
    double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (unsigned int b0 = 0; b0 < gridDim.x; b0 += 8) {
      double t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = (b0 + q < gridDim.x) ? __ldcg(a.partial + (int64_t)(b0 + q) * C::NM + e) : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] += t[q];
    }
    a.moments[e] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
  }
  if (tid == 0) *a.ticket = 0u;
}

// G[k,l] = moment(α_k + α_l); exponents of the library columns are passed by value
struct GatherArgs {
  const double* moments;
  double* gram;    // K×K
  double* header;  // out[0], out[1] (written if non-NULL)
  double n;
  int d, K, Q;
  unsigned char e[SB_MAX_TERMS * 3];
};

__device__ __forceinline__ int poly_count(int dt, int deg) {  // monomials of degree <= deg in dt variables
  if (deg < 0) return 0;
  if (dt == 0) return 1;
  if (dt == 1) return deg + 1;
  return (deg + 1) * (deg + 2) / 2;
}

__global__ void gram_gather_kernel(GatherArgs g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && g.header) { g.header[0] = 0.0; g.header[1] = g.n; }
  if (idx >= g.K * g.K) return;
  const int k = idx / g.K, l = idx % g.K;
  int al[3] = {0, 0, 0};
  for (int j = 0; j < g.d; ++j) al[j] = g.e[k * 3 + j] + g.e[l * 3 + j];
  const int a = al[0];
  const int dt = g.d - 1;
  int off = 0;
  for (int q = 0; q < a; ++q) off += poly_count(dt, g.Q - q);
  int rank = 0;  // graded order of the trailing exponents (b[,c]); within a degree block: by the last exponent
  if (dt == 1) rank = al[1];
  else if (dt == 2) { const int deg = al[1] + al[2]; rank = deg * (deg + 1) / 2 + al[2]; }
  g.gram[idx] = g.moments[off + rank];
}

template <int D, int P>
int run_moments(const float* x, int64_t n, const LibTab& t, double* gram_out, double* header, void* ws,
                int64_t ws_bytes, cudaStream_t s) {
  using C = MomCfg<D, P>;
  auto kern = moments_kernel<D, P>;
  static int grid_cached[64] = {0};
  int dev = 0;
  SB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return SB_ERR_INVALID; }
  static std::mutex init_mutex;   // the lazily built per-device launch geometry (the library is re-entrant)
  std::unique_lock<std::mutex> lock(init_mutex);
  if (grid_cached[dev] == 0) {
    SB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    int per_sm = 0, sms = 0;
    SB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, C::kSmemBytes));
    SB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm < 1) { set_error("moments kernel <%d,%d> does not fit on an SM", D, P); return SB_ERR_CUDA; }
    int g = per_sm * sms;
    if (g > kMaxPartialBlocks) g = kMaxPartialBlocks;
    grid_cached[dev] = g;
  }
  const int grid_max = grid_cached[dev];
  lock.unlock();
  MomArgs a{};
  a.x = x; a.n = n;
  a.n_bulk = n & ~(int64_t)3;
  a.n_tiles = (a.n_bulk + kTile - 1) / kTile;
  int64_t grid = (a.n_bulk + kTile / 2 - 1) / (kTile / 2);
  if (grid > grid_max) grid = grid_max;
  if (grid < 1) grid = 1;
  const int64_t need = kWsHeaderBytes + (grid + 1) * C::NM * (int64_t)sizeof(double);
  if (ws_bytes < need) {
    set_error("workspace too small: %lld < %lld bytes", (long long)ws_bytes, (long long)need);
    return SB_ERR_WORKSPACE;
  }
  a.ticket = reinterpret_cast<unsigned int*>(ws);
  a.partial = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kWsHeaderBytes);
  a.moments = a.partial + grid * C::NM;
  kern<<<(unsigned)grid, kThreads, C::kSmemBytes, s>>>(a);
  SB_LAUNCH_CHECK("moments_kernel");

  GatherArgs g{};
  g.moments = a.moments; g.gram = gram_out; g.header = header; g.n = (double)n;
  g.d = D; g.K = t.K; g.Q = 2 * P;
  for (int k = 0; k < t.K; ++k) {
    if (k == 0) continue;
    for (int j = 0; j < 3; ++j) g.e[k * 3 + j] = g.e[t.parent[k] * 3 + j];
    g.e[k * 3 + t.var[k]] += 1;
  }
  gram_gather_kernel<<<(t.K * t.K + 255) / 256, 256, 0, s>>>(g);
  SB_LAUNCH_CHECK("gram_gather_kernel");
  return SB_OK;
}

#define SB_MOMENT_SHAPES(X) X(2, 2) X(2, 3) X(3, 2) X(3, 3) X(3, 5)

}  // namespace

bool moments_supported(const LibTab& t) {
  if (t.sine || t.exp_) return false;
#define X(D, P) if (t.d == D && t.n_poly == n_poly_terms(D, P)) return true;
  SB_MOMENT_SHAPES(X)
#undef X
  return false;
}

int64_t moments_workspace_bytes(const LibTab& t) {
#define X(D, P) \
  if (t.d == D && t.n_poly == n_poly_terms(D, P)) \
    return kWsHeaderBytes + (int64_t)(kMaxPartialBlocks + 1) * MomCfg<D, P>::NM * (int64_t)sizeof(double);
  SB_MOMENT_SHAPES(X)
#undef X
  return 0;
}

int moments_gram(const float* x, int64_t n, const LibTab& t, double* gram_out, double* header, void* ws,
                 int64_t ws_bytes, cudaStream_t s) {
#define X(D, P) \
  if (t.d == D && t.n_poly == n_poly_terms(D, P)) return run_moments<D, P>(x, n, t, gram_out, header, ws, ws_bytes, s);
  SB_MOMENT_SHAPES(X)
#undef X
  set_error("no moment kernel for d=%d K=%d", t.d, t.K);
  return SB_ERR_UNSUPPORTED;
}

}  // namespace sb
