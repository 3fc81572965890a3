// sb_wsindy.cu — WSINDy weak-form integrals with trigonometric test functions.
//
// Reference: `sindy.py:332-347` builds V[j,t] = dt·sqrt(2/T)·sin((j+1)πt/T) and V'[j,t] as dense (n_test × T)
// fp32 matrices and `sindy.py:361-362` forms G = V·Θ(x), b = −V'·x with two GEMMs. Reading V would cost
// 8·n_test bytes per time sample against 4·d for x, so here the test functions are generated on the fly, in
// fp32 and with the reference's operation order (every intermediate rounded, no FMA contraction), and Θ is
// expanded in registers. blockIdx.x = test function j, blockIdx.y = trajectory: a thread accumulates the K
// entries G[j,:] and the d entries b[j,:] over its time samples; fp64 across threads; one block owns one
// output row, so there is no cross-block reduction and the result is deterministic.
#include "sb_common.cuh"

namespace sb {

namespace {

constexpr int kThreads = 256;

struct WsArgs {
  const float* x;   // n_traj × T × d
  int64_t T;
  float dt_f;       // fp32(dt): `self.dt = self.t[1] - self.t[0]`
  float tmax_f;     // fp32(t_max)
  float c1_f;       // fp32(sqrt(2 / t_max))
  int n_test;
  double* G;        // n_traj × n_test × K
  double* b;        // n_traj × n_test × d
};

template <int KMAX>
__global__ void __launch_bounds__(kThreads) wsindy_kernel(LibTab t, WsArgs a) {
  constexpr int NW = kThreads / 32;
  __shared__ float red[NW][KMAX + SB_MAX_DIM];
  const int d = t.d, K = t.K;
  const int j = blockIdx.x;
  const int64_t traj = blockIdx.y;
  const float* xt = a.x + traj * a.T * d;

  float m[KMAX], acc[KMAX], accb[SB_MAX_DIM], xv[SB_MAX_DIM];
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
#pragma unroll
  for (int q = 0; q < SB_MAX_DIM; ++q) accb[q] = 0.f;

  const float pi_f = 3.14159265358979323846f;       // fp32(torch.pi) == fp32(np.pi)
  const float kf = (float)(j + 1);
  const float kpi = __fmul_rn(kf, pi_f);             // k * pi
  // derivative prefactor: ((sqrt(2/T) * k) * pi) / T
  const float dpre = __fdiv_rn(__fmul_rn(__fmul_rn(a.c1_f, kf), pi_f), a.tmax_f);

  for (int64_t ti = threadIdx.x; ti < a.T; ti += kThreads) {
    const float tt = __fmul_rn((float)ti, a.dt_f);                  // torch.arange(n) * dt
    const float arg = __fdiv_rn(__fmul_rn(kpi, tt), a.tmax_f);      // k*pi*t / t_max
    float sn, cs;
    sincosf(arg, &sn, &cs);
    const float v = __fmul_rn(a.dt_f, __fmul_rn(a.c1_f, sn));       // V  = dt * (c1 * sin)
    const float vd = __fmul_rn(a.dt_f, __fmul_rn(dpre, cs));        // V' = dt * (dpre * cos)
#pragma unroll
    for (int q = 0; q < SB_MAX_DIM; ++q)
      if (q < d) xv[q] = __ldg(xt + ti * d + q);
    m[0] = 1.f;
    for (int q = 0; q < d; ++q) m[1 + q] = xv[q];
    for (int k = 1 + d; k < t.n_poly; ++k) m[k] = m[t.parent[k]] * xv[t.var[k]];
    int k = t.n_poly;
    if (t.sine) for (int q = 0; q < d; ++q) m[k++] = sinf(xv[q]);
    if (t.exp_) for (int q = 0; q < d; ++q) m[k++] = expf(xv[q]);
    for (int q = 0; q < K; ++q) acc[q] = fmaf(v, m[q], acc[q]);
#pragma unroll
    for (int q = 0; q < SB_MAX_DIM; ++q)
      if (q < d) accb[q] = fmaf(vd, xv[q], accb[q]);
  }

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int q = 0; q < K; ++q) {
    const float s = warp_sum(acc[q]);
    if (lane == 0) red[wid][q] = s;
  }
#pragma unroll
  for (int q = 0; q < SB_MAX_DIM; ++q) {
    if (q < d) {
      const float s = warp_sum(accb[q]);
      if (lane == 0) red[wid][K + q] = s;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K + d; e += kThreads) {
    double s = 0.0;
#pragma unroll
    for (int wq = 0; wq < NW; ++wq) s += (double)red[wq][e];
    if (e < K) a.G[(traj * a.n_test + j) * K + e] = s;
    else a.b[(traj * a.n_test + j) * d + (e - K)] = -s;
  }
}

}  // namespace

int wsindy_integrals(const float* x, int64_t n_traj, int64_t T, const LibTab& t, float dt, double t_max,
                     int n_test, double* G, double* b, cudaStream_t s) {
  if (n_traj > 65535) { set_error("n_traj=%lld > 65535 per call", (long long)n_traj); return SB_ERR_UNSUPPORTED; }
  WsArgs a{};
  a.x = x; a.T = T; a.dt_f = dt; a.tmax_f = (float)t_max; a.c1_f = (float)sqrt(2.0 / t_max);
  a.n_test = n_test; a.G = G; a.b = b;
  dim3 grid((unsigned)n_test, (unsigned)n_traj);
  if (t.K <= 16) wsindy_kernel<16><<<grid, kThreads, 0, s>>>(t, a);
  else if (t.K <= 64) wsindy_kernel<64><<<grid, kThreads, 0, s>>>(t, a);
  else wsindy_kernel<256><<<grid, kThreads, 0, s>>>(t, a);
  SB_LAUNCH_CHECK("wsindy_kernel");
  return SB_OK;
}

}  // namespace sb
