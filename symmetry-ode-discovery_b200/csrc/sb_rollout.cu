// sb_rollout.cu — lock-step fixed-step integration of dx/dt = Θ(x)·Wᵀ for a batch of initial conditions.
//
// Two reference functions share this kernel family:
//   * `model_utils.py:223-255` odeint(f, x0, t, dt, method, full_traj): torch fp32, Euler or classic RK4,
//     trajectory WITHOUT x0;
//   * `data_utils/ode.py:7-28` solve_ode_batch: NumPy float64 RK4 that also records dx = f(x) at every stored
//     row (row 0 = x0) and makes num_steps-1 updates.
// One thread integrates one initial condition for all steps: the state, the 4 stage derivatives and the K
// monomials stay in registers; W sits in the constant bank; the only HBM traffic is the strided trajectory
// store (and x0). The update formulas reproduce the reference's operation order with explicit
// round-to-nearest mul/add (no FMA contraction across the reference's separate tensor ops), so that fp32
// rollouts track torch's and fp64 rollouts track NumPy's to rounding.
// Register note (fp32, d=3, K=56): ptxas hoists the 168 loop-invariant coefficient loads out of the time loop into
// vector registers (164 registers, 3 blocks of 128 threads per SM). Keeping W in the constant bank instead (an opaque
// per-call-site offset defeats the hoisting: 52 registers, 9 blocks per SM) costs one indexed LDCU.64 per packed FMA
// and was measured SLOWER (73.1 ms against 67.9 for 1e6 ICs x 2000 steps): issue-bound. Inline `ld.const` with
// immediate addresses is hoisted by ptxas just the same.
#include <cstdlib>

#include "sb_common.cuh"

namespace sb {

namespace {

constexpr int kThreads = 128;

__constant__ double c_rw[kConstW];  // W as float or double (reinterpreted), 16 KB

template <class T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <class T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <class T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }

// ---- right-hand sides ---------------------------------------------------------------------------
// specialised: compile-time library (polynomial block + optional sin / exp columns), W from the constant bank
template <int D, int P, class T, int S = 0, int E = 0>
struct SpecRhs {
  static constexpr int DN = D;
  __device__ __forceinline__ int dim() const { return D; }
  // `off` (always 0, opaque, different per call site) only matters in fp64: it keeps ptxas from hoisting coefficient
  // loads out of the time loop into registers it does not have (168 registers + spills -> 136, none)
  __device__ __forceinline__ void eval(const T (&x)[D], T (&f)[D], int off = 0) const {
    constexpr int NP = Poly<D, P>::K;
    constexpr int K = NP + D * (S + E);
    T m[K];
    {
      T mp[NP];
      expand_poly<D, P>(x, mp);
      static_for<0, NP>([&](auto k) { m[k] = mp[k]; });
      if constexpr (S) static_for<0, D>([&](auto j) { m[NP + j] = sin(x[j]); });          // sinf / sin by overload
      if constexpr (E) static_for<0, D>([&](auto j) { m[NP + D * S + j] = exp(x[j]); });
    }
    if constexpr (std::is_same<T, float>::value && K % 2 == 0) {
      // fp32: packed fma.rn.f32x2 over adjacent columns, W pairs straight from the constant bank
      const float2* w2 = reinterpret_cast<const float2*>(c_rw);
      static_for<0, D>([&](auto ic) {
        constexpr int i = ic;
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
        static_for<0, K / 2>([&](auto kc) {
          constexpr int kk = kc;
          const float2 m2 = make_float2(m[2 * kk], m[2 * kk + 1]);
          if constexpr (kk % 2 == 0) s0 = __ffma2_rn(w2[i * (K / 2) + kk], m2, s0);
          else s1 = __ffma2_rn(w2[i * (K / 2) + kk], m2, s1);
        });
        f[i] = (s0.x + s0.y) + (s1.x + s1.y);
      });
      return;
    }
    const T* w = reinterpret_cast<const T*>(c_rw) + (std::is_same<T, double>::value ? 2 * off : 0);
    static_for<0, D>([&](auto ic) {
      constexpr int i = ic;
      // two partial sums (even / odd columns) keep the dependent chain short
      T s0 = T(0), s1 = T(0);
      static_for<0, K>([&](auto kc) {
        constexpr int k = kc;
        if constexpr (k % 2 == 0) s0 = fma(w[i * K + k], m[k], s0);
        else s1 = fma(w[i * K + k], m[k], s1);
      });
      f[i] = s0 + s1;
    });
  }
};

// generic: runtime table, W from global memory
template <class T, int KMAX>
struct GenRhs {
  static constexpr int DN = SB_MAX_DIM;
  LibTab t;
  const T* w;
  __device__ __forceinline__ int dim() const { return t.d; }
  __device__ __forceinline__ void eval(const T (&x)[SB_MAX_DIM], T (&f)[SB_MAX_DIM], int = 0) const {
    T m[KMAX];
    m[0] = T(1);
    for (int j = 0; j < t.d; ++j) m[1 + j] = x[j];
    for (int k = 1 + t.d; k < t.n_poly; ++k) m[k] = m[t.parent[k]] * x[t.var[k]];
    int k = t.n_poly;
    if (t.sine) for (int j = 0; j < t.d; ++j) m[k++] = sin(x[j]);
    if (t.exp_) for (int j = 0; j < t.d; ++j) m[k++] = exp(x[j]);
    for (int i = 0; i < t.d; ++i) {
      T s = T(0);
      for (int q = 0; q < t.K; ++q) s = fma(__ldg(w + i * t.K + q), m[q], s);
      f[i] = s;
    }
  }
};

struct RollArgs {
  const void* x0;
  int64_t n_ics;
  double dt;
  int64_t n_steps;
  int64_t stride;
  int method;
  int record_dx;
  long long opaque;   // always 0, unknown to the compiler: see eval_multi
  void* x_out;
  void* dx_out;
  void* x_last;
};

template <class RHS, class T>
__device__ __forceinline__ void rollout_body(const RHS& rhs, const RollArgs& a) {
  constexpr int DN = RHS::DN;
  const int d = rhs.dim();
  const int64_t ic = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= a.n_ics) return;
  const T* x0 = reinterpret_cast<const T*>(a.x0);
  T* xo = reinterpret_cast<T*>(a.x_out);
  T* dxo = reinterpret_cast<T*>(a.dx_out);
  T x[DN], k1[DN], k2[DN], k3[DN], k4[DN], xt[DN];
#pragma unroll
  for (int j = 0; j < DN; ++j) x[j] = (j < d) ? x0[ic * d + j] : T(0);

  const T dt = (T)a.dt;
  const int64_t row_elems = a.n_ics * d;
  auto store = [&](T* base, int64_t row, const T (&v)[DN]) {
#pragma unroll
    for (int j = 0; j < DN; ++j)
      if (j < d) base[row * row_elems + ic * d + j] = v[j];
  };

  if (a.record_dx) {
    // `data_utils/ode.py:14-26`: k_i = dt*f(.), x += (k1 + 2k2 + 2k3 + k4)/6, rows include x0
    const T half = T(0.5), two = T(2), six = T(6);
    int64_t row = 0, until = 0;   // next stored row and steps until it (no 64-bit division in the step loop)
    for (int64_t i = 0; i < a.n_steps; ++i) {
      rhs.eval(x, k1, (int)(i & a.opaque));
      if (until == 0) {
        if (xo) store(xo, row, x);
        if (dxo) store(dxo, row, k1);
        ++row;
        until = a.stride;
      }
      --until;
      if (i == a.n_steps - 1) break;
      if (a.method == SB_EULER) {
#pragma unroll
        for (int j = 0; j < DN; ++j) x[j] = add_rn(x[j], mul_rn(dt, k1[j]));
        continue;
      }
#pragma unroll
      for (int j = 0; j < DN; ++j) { k1[j] = mul_rn(dt, k1[j]); xt[j] = add_rn(x[j], mul_rn(half, k1[j])); }
      rhs.eval(xt, k2, (int)((i + 1) & a.opaque));
#pragma unroll
      for (int j = 0; j < DN; ++j) { k2[j] = mul_rn(dt, k2[j]); xt[j] = add_rn(x[j], mul_rn(half, k2[j])); }
      rhs.eval(xt, k3, (int)((i + 2) & a.opaque));
#pragma unroll
      for (int j = 0; j < DN; ++j) { k3[j] = mul_rn(dt, k3[j]); xt[j] = add_rn(x[j], k3[j]); }
      rhs.eval(xt, k4, (int)((i + 3) & a.opaque));
#pragma unroll
      for (int j = 0; j < DN; ++j) {
        k4[j] = mul_rn(dt, k4[j]);
        T s = add_rn(k1[j], mul_rn(two, k2[j]));
        s = add_rn(s, mul_rn(two, k3[j]));
        s = add_rn(s, k4[j]);
        x[j] = add_rn(x[j], div_rn(s, six));
      }
    }
  } else {
    // `model_utils.py:236-253`: x0 + dt/2*k1 etc.; python scalars dt/2, dt/6 are rounded to T once
    const T hdt = (T)(a.dt / 2.0), sdt = (T)(a.dt / 6.0), two = T(2);
    int64_t row = 0, until = a.stride;   // next stored row and steps until it (no 64-bit division in the step loop)
    for (int64_t s = 1; s <= a.n_steps; ++s) {
      rhs.eval(x, k1, (int)(s & a.opaque));
      if (a.method == SB_EULER) {
#pragma unroll
        for (int j = 0; j < DN; ++j) x[j] = add_rn(x[j], mul_rn(dt, k1[j]));
      } else {
#pragma unroll
        for (int j = 0; j < DN; ++j) xt[j] = add_rn(x[j], mul_rn(hdt, k1[j]));
        rhs.eval(xt, k2, (int)((s + 1) & a.opaque));
#pragma unroll
        for (int j = 0; j < DN; ++j) xt[j] = add_rn(x[j], mul_rn(hdt, k2[j]));
        rhs.eval(xt, k3, (int)((s + 2) & a.opaque));
#pragma unroll
        for (int j = 0; j < DN; ++j) xt[j] = add_rn(x[j], mul_rn(dt, k3[j]));
        rhs.eval(xt, k4, (int)((s + 3) & a.opaque));
#pragma unroll
        for (int j = 0; j < DN; ++j) {
          T acc = add_rn(k1[j], mul_rn(two, k2[j]));
          acc = add_rn(acc, mul_rn(two, k3[j]));
          acc = add_rn(acc, k4[j]);
          x[j] = add_rn(x[j], mul_rn(sdt, acc));
        }
      }
      if (--until == 0) {
        if (xo) store(xo, row, x);
        ++row;
        until = a.stride;
      }
    }
  }
  if (a.x_last) {
    T* xl = reinterpret_cast<T*>(a.x_last);
#pragma unroll
    for (int j = 0; j < DN; ++j)
      if (j < d) xl[ic * d + j] = x[j];
  }
}

template <int D, int P, class T, int S = 0, int E = 0>
__global__ void __launch_bounds__(kThreads) rollout_spec_kernel(RollArgs a) {
  SpecRhs<D, P, T, S, E> rhs;
  rollout_body<SpecRhs<D, P, T, S, E>, T>(rhs, a);
}

template <class T, int KMAX>
__global__ void __launch_bounds__(kThreads) rollout_gen_kernel(LibTab t, const T* w, RollArgs a) {
  GenRhs<T, KMAX> rhs{t, w};
  rollout_body<GenRhs<T, KMAX>, T>(rhs, a);
}

// ---- fp32 RK4 (odeint layout), NB initial conditions per thread --------------------------------------------------
// The coefficient pairs are loop-invariant, so ptxas either keeps all of them in vector registers (the kernel above:
// 164 registers, packed FMAs with three vector operands) or — when an opaque offset `off` (always 0, different per
// call site) defeats the hoisting — reloads every pair with one indexed LDCU.64 per packed FMA, which is issue-bound
// with one initial condition per thread. With NB = 2 every loaded pair feeds NB packed FMAs: W stays in the constant
// bank (uniform-register operands), the thread needs ~100 registers and carries two independent dependency chains.
// Same operation order per initial condition as rollout_body's odeint branch: bit-identical trajectories.
template <int D, int P, int NB>
__device__ __forceinline__ void eval_multi(const float (&x)[NB][D], float (&f)[NB][D], int off) {
  constexpr int K = Poly<D, P>::K;
  using L = Poly<D, P>;
  static_assert(K % 2 == 0, "packed pairs");
  const float2* w2 = reinterpret_cast<const float2*>(c_rw) + off;
  float m[NB][K];
  float2 s0[NB][D], s1[NB][D];
  static_for<0, NB>([&](auto b) {
    static_for<0, D>([&](auto i) { s0[b][i] = make_float2(0.f, 0.f); s1[b][i] = make_float2(0.f, 0.f); });
  });
  static_for<0, K>([&](auto kc) {
    constexpr int k = kc;
    static_for<0, NB>([&](auto bc) {
      constexpr int b = bc;
      if constexpr (k == 0) m[b][0] = 1.f;
      else if constexpr (k <= D) m[b][k] = x[b][k - 1];
      else {
        constexpr int pk = L::tab.parent[k];
        constexpr int vk = L::tab.var[k];
        m[b][k] = m[b][pk] * x[b][vk];
      }
    });
    if constexpr (k % 2 == 1) {
      constexpr int kk = k / 2;
      static_for<0, D>([&](auto ic) {
        constexpr int i = ic;
        const float2 w = w2[i * (K / 2) + kk];
        static_for<0, NB>([&](auto bc) {
          constexpr int b = bc;
          const float2 m2 = make_float2(m[b][k - 1], m[b][k]);
          if constexpr (kk % 2 == 0) s0[b][i] = __ffma2_rn(w, m2, s0[b][i]);
          else s1[b][i] = __ffma2_rn(w, m2, s1[b][i]);
        });
      });
    }
  });
  static_for<0, NB>([&](auto b) {
    static_for<0, D>([&](auto i) { f[b][i] = (s0[b][i].x + s0[b][i].y) + (s1[b][i].x + s1[b][i].y); });
  });
}

// Blocks of kMultiThreads = 64 threads (128 initial conditions): with the whole batch resident at once — 1e6 / 8 ICs per
// GPU in the sharded C5 rollout is two thirds of one wave — the SM loads differ by at most one block, and the run time is
// set by the fullest SM: 128-thread blocks gave 489 blocks over 148 SMs = 3 or 4 per SM (+21 % on the slowest, measured
// as 6.58x at 8 GPUs instead of 8x); 64-thread blocks give 6 or 7.
constexpr int kMultiThreads = 64;

template <int D, int P, int NB>
__global__ void __launch_bounds__(kMultiThreads) rollout_rk4_multi_kernel(RollArgs a) {
  constexpr int kThreads = kMultiThreads;
  const int64_t base = (int64_t)blockIdx.x * (kThreads * NB) + threadIdx.x;
  if (base >= a.n_ics) return;
  const float* x0 = reinterpret_cast<const float*>(a.x0);
  float* xo = reinterpret_cast<float*>(a.x_out);
  float x[NB][D], k1[NB][D], k2[NB][D], k3[NB][D], k4[NB][D], xt[NB][D];
  bool live[NB];
  static_for<0, NB>([&](auto bc) {
    constexpr int b = bc;
    const int64_t ic = base + (int64_t)b * kThreads;
    live[b] = ic < a.n_ics;
    static_for<0, D>([&](auto j) { x[b][j] = live[b] ? x0[ic * D + j] : 0.f; });
  });
  const float dt = (float)a.dt, hdt = (float)(a.dt / 2.0), sdt = (float)(a.dt / 6.0), two = 2.f;
  const int64_t row_elems = a.n_ics * D;
  int64_t row = 0, until = a.stride;
  for (int64_t s = 1; s <= a.n_steps; ++s) {
    eval_multi<D, P, NB>(x, k1, 2 * (int)(s & a.opaque));
    static_for<0, NB>([&](auto b) { static_for<0, D>([&](auto j) { xt[b][j] = __fadd_rn(x[b][j], __fmul_rn(hdt, k1[b][j])); }); });
    eval_multi<D, P, NB>(xt, k2, 2 * (int)((s + 1) & a.opaque));
    static_for<0, NB>([&](auto b) { static_for<0, D>([&](auto j) { xt[b][j] = __fadd_rn(x[b][j], __fmul_rn(hdt, k2[b][j])); }); });
    eval_multi<D, P, NB>(xt, k3, 2 * (int)((s + 2) & a.opaque));
    static_for<0, NB>([&](auto b) { static_for<0, D>([&](auto j) { xt[b][j] = __fadd_rn(x[b][j], __fmul_rn(dt, k3[b][j])); }); });
    eval_multi<D, P, NB>(xt, k4, 2 * (int)((s + 3) & a.opaque));
    static_for<0, NB>([&](auto b) {
      static_for<0, D>([&](auto j) {
        float acc = __fadd_rn(k1[b][j], __fmul_rn(two, k2[b][j]));
        acc = __fadd_rn(acc, __fmul_rn(two, k3[b][j]));
        acc = __fadd_rn(acc, k4[b][j]);
        x[b][j] = __fadd_rn(x[b][j], __fmul_rn(sdt, acc));
      });
    });
    if (--until == 0) {
      if (xo) {
        static_for<0, NB>([&](auto bc) {
          constexpr int b = bc;
          const int64_t ic = base + (int64_t)b * kThreads;
          if (live[b]) static_for<0, D>([&](auto j) { xo[row * row_elems + ic * D + j] = x[b][j]; });
        });
      }
      ++row;
      until = a.stride;
    }
  }
  if (a.x_last) {
    float* xl = reinterpret_cast<float*>(a.x_last);
    static_for<0, NB>([&](auto bc) {
      constexpr int b = bc;
      const int64_t ic = base + (int64_t)b * kThreads;
      if (live[b]) static_for<0, D>([&](auto j) { xl[ic * D + j] = x[b][j]; });
    });
  }
}

#define SB_ROLL_SHAPES(X) X(2, 2) X(2, 3) X(3, 2) X(3, 3) X(3, 5)

// SB_ROLLOUT_MULTI=0 selects the one-initial-condition-per-thread kernel for the fp32 RK4 rollout (A/B, bitwise test)
bool rollout_multi_enabled() {
  const char* e = getenv("SB_ROLLOUT_MULTI");
  return !(e && e[0] == '0');
}

template <class T>
int launch_rollout(const LibTab& t, const void* w, const RollArgs& a, cudaStream_t s) {
  const unsigned grid = (unsigned)((a.n_ics + kThreads - 1) / kThreads);
  if (!t.sine && !t.exp_) {
#define X(D, P)                                                                                              \
  if (t.d == D && t.n_poly == n_poly_terms(D, P)) {                                                          \
    SB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_rw, w, sizeof(T) * D * Poly<D, P>::K, 0, cudaMemcpyDeviceToDevice, s)); \
    if constexpr (std::is_same<T, float>::value && Poly<D, P>::K % 2 == 0) {                                   \
      if (a.method == SB_RK4 && !a.record_dx && rollout_multi_enabled()) {                                     \
        const unsigned g2 = (unsigned)((a.n_ics + 2 * kMultiThreads - 1) / (2 * kMultiThreads));               \
        rollout_rk4_multi_kernel<D, P, 2><<<g2, kMultiThreads, 0, s>>>(a);                                     \
        SB_LAUNCH_CHECK("rollout_rk4_multi_kernel");                                                           \
        return SB_OK;                                                                                          \
      }                                                                                                        \
    }                                                                                                          \
    rollout_spec_kernel<D, P, T><<<grid, kThreads, 0, s>>>(a);                                                \
    SB_LAUNCH_CHECK("rollout_spec_kernel");                                                                  \
    return SB_OK;                                                                                            \
  }
    SB_ROLL_SHAPES(X)
#undef X
  }
  // Lotka-Volterra in canonical coordinates (`data_utils/lotka.py:33-41`; config 3's library): (2, 2) + exp, K = 8
  if (t.d == 2 && t.n_poly == n_poly_terms(2, 2) && !t.sine && t.exp_) {
    SB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_rw, w, sizeof(T) * 2 * 8, 0, cudaMemcpyDeviceToDevice, s));
    rollout_spec_kernel<2, 2, T, 0, 1><<<grid, kThreads, 0, s>>>(a);
    SB_LAUNCH_CHECK("rollout_spec_kernel");
    return SB_OK;
  }
  if (t.K <= 16) rollout_gen_kernel<T, 16><<<grid, kThreads, 0, s>>>(t, reinterpret_cast<const T*>(w), a);
  else if (t.K <= 64) rollout_gen_kernel<T, 64><<<grid, kThreads, 0, s>>>(t, reinterpret_cast<const T*>(w), a);
  else rollout_gen_kernel<T, 256><<<grid, kThreads, 0, s>>>(t, reinterpret_cast<const T*>(w), a);
  SB_LAUNCH_CHECK("rollout_gen_kernel");
  return SB_OK;
}

}  // namespace

int rollout(const void* x0, int64_t n_ics, const LibTab& t, const void* w, double dt, int64_t n_steps,
            int64_t stride, int method, int dtype, int record_dx, void* x_out, void* dx_out, void* x_last,
            cudaStream_t s) {
  RollArgs a{};
  a.x0 = x0; a.n_ics = n_ics; a.dt = dt; a.n_steps = n_steps; a.stride = stride; a.method = method;
  a.record_dx = record_dx; a.opaque = 0; a.x_out = x_out; a.dx_out = dx_out; a.x_last = x_last;
  if (dtype == SB_F32) return launch_rollout<float>(t, w, a, s);
  return launch_rollout<double>(t, w, a, s);
}

}  // namespace sb
