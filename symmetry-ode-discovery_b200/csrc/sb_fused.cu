// sb_fused.cu — register-resident fused SINDy train step for the hot polynomial libraries.
//
// Replaces, in ONE pass over (x, dx), the reference's `regressor(x)` (Θ materialised column by column,
// `sindy.py:79-82`), `MSELoss` (`train.py:663-664`) and the Θ-side of `loss.backward()` (`train.py:689`):
//   r = Θ(x)·Wᵀ − dx,   out = { Σ r², Σ_n r_i Θ_k }   [+ loss and dL/dΞ when the closure epilogue is requested,
//   + the optimiser update of `train.py:530` and the next launch's coefficients for sb_fit_step: a whole training
//   iteration is then this one kernel, across GPUs too (all-reduce over NVLink peer memory inside the last block)].
// Θ never exists in memory: each thread expands the K monomials of its sample in registers by the
// parent*variable recurrence, forms the d predictions, and accumulates the d×K outer product r ⊗ Θ into
// private fp32 accumulators. All d·K FMAs of the prediction and of the gradient are issued as packed
// `fma.rn.f32x2` (SASS FFMA2: two fp32 FMAs per lane per issue slot) over adjacent library columns; W comes
// from the constant bank (uniform datapath), so the only vector-register operands are Θ pairs, r and the
// accumulators.
//
// Data movement: x and dx tiles are streamed HBM -> shared memory with 1-D TMA bulk copies
// (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) into a kStages-deep ring. A "full" mbarrier per stage
// carries the transaction bytes; an "empty" mbarrier per stage counts one arrival per warp, so the elected
// producer thread refills a stage without any CTA-wide barrier in the steady state. X and dX are read exactly
// once, 8·d bytes per sample. Grid = resident CTAs (multiple of the SM count), one contiguous sample range per CTA
// (equal to within 4 samples), one fp32 partial row per CTA, and an ordered, error-free (TwoSum) reduction of the rows
// by the last block, which stages them in the idle tile ring with TMA => deterministic, fp64-equivalent sums.
#include <cstdlib>

#include <atomic>
#include <mutex>

#include "sb_common.cuh"
#include "sb_tma.cuh"

namespace sb {

namespace {

// W = Ξ⊙mask lives in the constant bank so that the kernels read it through the uniform datapath (LDCU -> UR operands).
// The bank is cut in kWSlots slots of kWSlotFloats floats (every specialised library needs <= 2·d·K2 = 168):
//   slot 0      scratch of the stateless calls (sb_train_step, sb_closure[_peer], sb_forward, sb_backward): each packs its
//               W right before its kernel ON ITS STREAM, so calls that are stream-ordered with each other never see a
//               foreign W; a call arriving on another stream first waits for the previous user (guard_scratch below);
//   slot 1      resident coefficients of the one-launch iteration (sb_fit_step / sb_load_w): the kernel's epilogue writes
//               the next Ξ⊙mask back into it and no stateless call touches it — a validation forward or an STLSQ closure
//               between two iterations cannot clobber it. Two FITS that alternate on one device share it: the host side
//               keeps a generation count per device and re-packs when the slot was last loaded by somebody else
//               (native.load_w / FitStepper._own_slot).
// The slot is a TEMPLATE parameter of the kernel, not an argument: with an immediate address ptxas fetches W with
// LDCU.128 (27 per sample at d = 3, K = 56); a uniform-register offset makes it fall back to 84 LDCU.64.
constexpr int kWSlots = 2;
constexpr int kWSlotFloats = 512;                      // 2 KB per slot (every specialised library needs <= 168 floats)
constexpr int kWSlotPairs = kWSlotFloats / 2;
static_assert(kWSlots * kWSlotFloats <= kConstW, "slots exceed the constant array");
__constant__ float2 c_w2[kConstW / 2];

unsigned long long* g_trace = nullptr;   // set by sb_debug_trace (timeline of the fused kernel's phases)
__device__ __forceinline__ void stamp(unsigned long long* trace, int slot) {
  if (trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    trace[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

// ---- configuration per library ------------------------------------------------------------------
// VAR selects a tuning variant (A/B-tested on the B200):
//   bit 0: 1 = per-stage "empty" mbarriers (no __syncthreads in the tile loop), 0 = CTA barrier per tile
//   bit 1: 1 = two partial sums per equation in the prediction, 0 = one
//   bit 2: 1 = transposed-butterfly CTA reduction (~NV shuffles), 0 = one 5-step shuffle reduction per value
//   bits 3-4: tile ring: 0 = 1024 samples x 4 stages, 1 = 2048 x 3, 2 = 4096 x 2, 3 = 2048 x 4
//   bit 6: 1 = refill the stage consumed TWO tiles ago (its empty barrier has long completed), 0 = one tile ago
//   bit 5: 1 = the next sample's x/dx are read from shared memory before the current sample is expanded
//   bit 7: 1 = equation-by-equation order (pred_i, r_i, acc_i += r_i·Θ for i = 0..d-1) so that the independent
//          accumulation of equation i can fill the dependent prediction chains of equation i+1
template <int D, int P, int VAR = 1>
struct Cfg {
  static constexpr int K = LibCode<D, P>::K;   // P is a library CODE (sb_common.cuh): degree, +10 exp, +20 sine
  static constexpr int K2 = (K + 1) / 2;     // packed column pairs
  static constexpr int NV = D * K + 1;       // values reduced per CTA (grad + loss)
  static constexpr int kRow = (NV + 3) & ~3; // floats per partial row in the workspace (rows stay 16-byte aligned)
  static constexpr int kThreads = 256;
  static constexpr int kWarps = kThreads / 32;
  static constexpr int kTileCode = (VAR >> 3) & 3;        // samples per stage / ring depth
  static constexpr int kTile = kTileCode == 0 ? 1024 : (kTileCode == 2 ? 4096 : 2048);
  static constexpr int kStages = kTileCode == 0 ? 4 : (kTileCode == 1 ? 3 : (kTileCode == 2 ? 2 : 4));
  static constexpr bool kPrefetch = (VAR & 32) != 0;
  static constexpr int kLag = (VAR & 64) ? 2 : 1;          // the producer refills the stage consumed kLag tiles ago
  static constexpr bool kEmptyBarriers = (VAR & 1) != 0;
  static constexpr int kChains = (VAR & 2) ? 2 : 1;
  static constexpr bool kFoldReduce = (VAR & 4) != 0;
  static constexpr bool kEqOrder = (VAR & 128) != 0;
  // accumulators dominate the register budget: D*K2*2 of them
  static constexpr size_t kSmemData = (size_t)kStages * 2 * kTile * D * sizeof(float);
  // resident CTAs per SM the register budget is sized for: the accumulators (D*K2*2) dominate; a ring that fills more
  // than half of the shared memory leaves room for one CTA only, which may then use the whole register file
  static constexpr int kMinBlocks = (D * K2 * 2 + K > 150 || kSmemData > 110 * 1024) ? 1
                                    : ((D * K2 * 2 + K > 40 || kSmemData > 72 * 1024) ? 2 : 3);
  static constexpr size_t kSmemBytes = kSmemData + (2 * kStages + 1) * sizeof(uint64_t) + 16;
};

enum { LEFT_RESIDUAL = 0, LEFT_DX = 1 };

// one sample: expand Θ, predict, accumulate r ⊗ Θ
template <int D, int P, int LEFT, int VAR, int SLOT>
__device__ __forceinline__ void accumulate_sample(const float (&xs)[D], const float (&ds)[D],
                                                  float2 (&acc)[D][Cfg<D, P, VAR>::K2], float& lacc) {
  using C = Cfg<D, P, VAR>;
  constexpr int w_off = SLOT * kWSlotPairs;
  float m[C::K];
  expand_lib<D, P>(xs, m);
  float2 m2[C::K2];
  static_for<0, C::K2>([&](auto kc) {
    constexpr int kk = kc;
    m2[kk] = make_float2(m[2 * kk], (2 * kk + 1 < C::K) ? m[2 * kk + 1] : 0.f);
  });
  float r[D];
  if constexpr (LEFT == LEFT_RESIDUAL && C::kEqOrder) {
    constexpr int NC = 4;   // one equation at a time: four partial sums keep its chain short
    static_for<0, D>([&](auto ic) {
      constexpr int i = ic;
      float2 pred[NC];
      static_for<0, NC>([&](auto c) { pred[c] = make_float2(0.f, 0.f); });
      static_for<0, C::K2>([&](auto kc) {
        constexpr int kk = kc;
        pred[kk % NC] = __ffma2_rn(c_w2[w_off + i * C::K2 + kk], m2[kk], pred[kk % NC]);
      });
      const float s = ((pred[0].x + pred[0].y) + (pred[1].x + pred[1].y)) + ((pred[2].x + pred[2].y) + (pred[3].x + pred[3].y));
      r[i] = s - ds[i];
      lacc = fmaf(r[i], r[i], lacc);
      const float2 r2 = make_float2(r[i], r[i]);
      static_for<0, C::K2>([&](auto kc) {
        constexpr int kk = kc;
        acc[i][kk] = __ffma2_rn(r2, m2[kk], acc[i][kk]);
      });
    });
    return;
  } else if constexpr (LEFT == LEFT_RESIDUAL) {
    constexpr int NC = C::kChains;
    float2 pred[D][NC];
    static_for<0, D>([&](auto i) { static_for<0, NC>([&](auto c) { pred[i][c] = make_float2(0.f, 0.f); }); });
    static_for<0, C::K2>([&](auto kc) {
      constexpr int kk = kc;
      static_for<0, D>([&](auto ic) {
        constexpr int i = ic;
        pred[i][kk % NC] = __ffma2_rn(c_w2[w_off + i * C::K2 + kk], m2[kk], pred[i][kk % NC]);
      });
    });
    static_for<0, D>([&](auto i) {
      float s = pred[i][0].x + pred[i][0].y;
      static_for<1, NC>([&](auto c) { s += pred[i][c].x + pred[i][c].y; });
      r[i] = s - ds[i];
      lacc = fmaf(r[i], r[i], lacc);
    });
  } else {
    static_for<0, D>([&](auto i) { r[i] = ds[i]; });
  }
  static_for<0, D>([&](auto ic) {
    constexpr int i = ic;
    const float2 r2 = make_float2(r[i], r[i]);
    static_for<0, C::K2>([&](auto kc) {
      constexpr int kk = kc;
      acc[i][kk] = __ffma2_rn(r2, m2[kk], acc[i][kk]);
    });
  });
}

// b^t by repeated squaring (≤ 32 multiplications; relative error ~1e-15, against `beta ** step` of torch's Adam)
__device__ __forceinline__ double powi(double b, unsigned int t) {
  double r = 1.0;
  while (t) {
    if (t & 1u) r *= b;
    b *= b;
    t >>= 1;
  }
  return r;
}

struct FusedArgs {
  const float* x;
  const float* dx;
  int64_t n;          // samples
  int64_t n_bulk;     // samples covered by TMA tiles (multiple of 4)
  int64_t n_tiles;
  float* partial;     // [grid][kRow]
  unsigned int* ticket;
  double* out;        // packed output base (may be NULL when only the closure epilogue is wanted)
  int64_t out_off;    // offset of this section's d×K block
  int out_transposed; // 0: [i*K+k], 1: [k*d+i]  (b section)
  int write_header;   // write out[0] (loss) and out[1] (n)
  // optional closure epilogue (single-rank): loss = Σr²/(n d) + w_l1‖Ξ‖₁, grad = 2/(n d)·Σ r⊗Θ ⊙ mask + w_l1 sign(Ξ)
  const float* xi;
  const float* mask;
  double w_l1;
  double w_mse;       // weight of the MSE term (w_sindy_x; 1 for sb_closure)
  float* loss_out;
  float* grad_out;
  // optional optimiser update fused into that epilogue (sb_fit_step): Ξ is advanced in place and Ξ⊙mask is packed
  // into the constant bank for the NEXT launch, so a whole training iteration is this one kernel
  FitArgs fit;
  float* w_const;     // device (global-alias) address of THIS call's constant slot
  // optional linear Lie-derivative regulariser as a quadratic form of w = vec(Ξ⊙mask): w_sym·wᵀHw (sb_fit_options)
  const float* sym_H; // (d·K)² fp32, symmetric, row-major
  double w_sym;
  unsigned long long* trace;  // debug: 16 globaltimer stamps per CTA (sb_debug_trace), NULL in production
  // optional in-kernel all-reduce over peer memory (NVLink P2P stores + flags): see the final phase
  PeerArgs peer;
};

template <int D, int P, int LEFT, int VAR, int SLOT>
__global__ void __launch_bounds__(Cfg<D, P, VAR>::kThreads, Cfg<D, P, VAR>::kMinBlocks)
fused_step_kernel(FusedArgs a) {
  using C = Cfg<D, P, VAR>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + C::kSmemData);
  uint64_t* empty = full + C::kStages;
  uint64_t* fin_bar = empty + C::kStages;   // completes when the last block's copy of the partial rows has landed
  __shared__ float red[C::kWarps][C::NV];
  __shared__ double fin[C::NV + 2];
  __shared__ int is_last;

  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  constexpr int kTileFloats = C::kTile * D;

  // Every CTA owns one contiguous range of samples, cut in quads (4 samples = 16·d bytes keep the TMA source
  // 16-byte aligned) so that all CTAs get the same amount of work to within 4 samples — no tile-count quantisation
  // in the tail, whatever the shard size. The range is consumed in tiles of kTile samples (the last one partial).
  const int64_t quads = a.n_bulk >> 2;
  const int64_t s_begin = 4 * (quads * (int64_t)blockIdx.x / (int64_t)gridDim.x);
  const int64_t s_end = 4 * (quads * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x);
  const int my_tiles = (int)((s_end - s_begin + C::kTile - 1) / C::kTile);
  auto tile_count = [&](int t) -> int {
    const int64_t rem = s_end - s_begin - (int64_t)t * C::kTile;
    return (int)(rem < C::kTile ? rem : C::kTile);
  };
  auto issue = [&](int t, int stage) {
    const int cnt = tile_count(t);
    const uint32_t bytes = (uint32_t)cnt * D * sizeof(float);
    float* sx = tiles + (size_t)stage * 2 * kTileFloats;
    const int64_t off = (s_begin + (int64_t)t * C::kTile) * D;
    mbar_expect_tx(&full[stage], 2 * bytes);
    tma_load_1d(sx, a.x + off, bytes, &full[stage]);
    tma_load_1d(sx + kTileFloats, a.dx + off, bytes, &full[stage]);
  };

  stamp(a.trace, 0);
  if (a.trace && tid == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    a.trace[(size_t)blockIdx.x * 16 + 7] = smid;
  }
  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], C::kWarps); }
    mbar_init(fin_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // the first tile travels alone (issuing the whole ring at once makes every CTA's first tile queue behind 148 × the
  // ring in the memory system: first data after 3.5 us instead of ~1 us); the rest of the ring follows once it landed
  if (tid == 0 && my_tiles > 0) issue(0, 0);

  float2 acc[D][C::K2];
  static_for<0, D>([&](auto i) {
    static_for<0, C::K2>([&](auto kk) { acc[i][kk] = make_float2(0.f, 0.f); });
  });
  float lacc = 0.f;

  for (int it = 0; it < my_tiles; ++it) {
    const int stage = it % C::kStages;
    const uint32_t parity = (uint32_t)(it / C::kStages) & 1u;
    if constexpr (C::kEmptyBarriers) {
      // refill the stage consumed kLag iterations ago: by now every warp has (almost surely) released it
      if (tid == 0 && it >= C::kLag) {
        const int next = it + C::kStages - C::kLag;
        if (next < my_tiles) {
          const int ps = (it - C::kLag) % C::kStages;
          mbar_wait(&empty[ps], (uint32_t)((it - C::kLag) / C::kStages) & 1u);
          issue(next, ps);
        }
      }
    }
    mbar_wait(&full[stage], parity);
    if (it == 0) {
      if (tid == 0) {
        for (int s = 1; s < C::kStages; ++s)
          if (s < my_tiles) issue(s, s);
      }
      stamp(a.trace, 1);
    }
    const float* sx = tiles + (size_t)stage * 2 * kTileFloats;
    const float* sd = sx + kTileFloats;
    const int cnt = tile_count(it);
    if constexpr (C::kPrefetch) {
      float xn[D], dn[D];
      if (tid < cnt) static_for<0, D>([&](auto q) { xn[q] = sx[tid * D + q]; dn[q] = sd[tid * D + q]; });
#pragma unroll 1
      for (int j = tid; j < cnt; j += C::kThreads) {
        float xs[D], ds[D];
        static_for<0, D>([&](auto q) { xs[q] = xn[q]; ds[q] = dn[q]; });
        const int jn = j + C::kThreads;
        if (jn < cnt) static_for<0, D>([&](auto q) { xn[q] = sx[jn * D + q]; dn[q] = sd[jn * D + q]; });
        accumulate_sample<D, P, LEFT, VAR, SLOT>(xs, ds, acc, lacc);
      }
    } else {
#pragma unroll 1
      for (int j = tid; j < cnt; j += C::kThreads) {
        float xs[D], ds[D];
        static_for<0, D>([&](auto q) { xs[q] = sx[j * D + q]; ds[q] = sd[j * D + q]; });
        accumulate_sample<D, P, LEFT, VAR, SLOT>(xs, ds, acc, lacc);
      }
    }
    if constexpr (C::kEmptyBarriers) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);   // this warp is done with the stage
    } else {
      __syncthreads();  // every thread is done with this stage before it is refilled
      if (tid == 0) {
        const int next = it + C::kStages;
        if (next < my_tiles) issue(next, stage);
      }
    }
  }

  // ragged tail (n % 4 samples) straight from global memory, by the first threads of block 0
  if (blockIdx.x == 0) {
    const int64_t j = a.n_bulk + tid;
    if (j < a.n) {
      float xs[D], ds[D];
      static_for<0, D>([&](auto q) { xs[q] = __ldg(a.x + j * D + q); ds[q] = __ldg(a.dx + j * D + q); });
      accumulate_sample<D, P, LEFT, VAR, SLOT>(xs, ds, acc, lacc);
    }
  }

  stamp(a.trace, 2);
  // ---- CTA reduction within the warp (fp32), fp64 across warps below ----
  if constexpr (C::kFoldReduce) {
    constexpr int V = ((C::NV + 31) / 32) * 32;
    float v[V];
    static_for<0, D>([&](auto ic) {
      constexpr int i = ic;
      static_for<0, C::K2>([&](auto kc) {
        constexpr int kk = kc;
        v[i * C::K + 2 * kk] = acc[i][kk].x;
        if constexpr (2 * kk + 1 < C::K) v[i * C::K + 2 * kk + 1] = acc[i][kk].y;
      });
    });
    v[C::NV - 1] = lacc;
    static_for<C::NV, V>([&](auto i) { v[i] = 0.f; });
    warp_fold<V>(v, lane);
    static_for<0, V / 32>([&](auto ic) {
      constexpr int i = ic;
      const int e = warp_fold_index<V>(i, lane);
      if (e < C::NV) red[wid][e] = v[i];
    });
  } else {
    static_for<0, D>([&](auto ic) {
      constexpr int i = ic;
      static_for<0, C::K2>([&](auto kc) {
        constexpr int kk = kc;
        const float vx = warp_sum(acc[i][kk].x);
        const float vy = warp_sum(acc[i][kk].y);
        if (lane == 0) {
          red[wid][i * C::K + 2 * kk] = vx;
          if (2 * kk + 1 < C::K) red[wid][i * C::K + 2 * kk + 1] = vy;
        }
      });
    });
    const float vl = warp_sum(lacc);
    if (lane == 0) red[wid][C::NV - 1] = vl;
  }
  __syncthreads();
  // per-CTA partial row: the 8 warp sums are added in fp64 and stored as fp32 — the same precision class as the
  // per-thread fp32 accumulators they come from; everything across CTAs (and ranks) is fp64 again
  float* mine = a.partial + (int64_t)blockIdx.x * C::kRow;
  for (int e = tid; e < C::kRow; e += C::kThreads) {
    double v = 0.0;
    if (e < C::NV) {
#pragma unroll
      for (int wq = 0; wq < C::kWarps; ++wq) v += (double)red[wq][e];
    }
    mine[e] = (float)v;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();   // cumulative: orders the whole CTA's row (observed through the barrier) before the ticket
    is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1u);
  }
  __syncthreads();
  stamp(a.trace, 3);
  if (!is_last) return;
  __threadfence();

  // the last block's epilogue inputs travel while it reduces (one element per thread: d·K <= 256 for every shape)
  static_assert(D * C::K <= C::kThreads, "epilogue assumes one coefficient per thread");
  float ep_xi = 0.f, ep_mk = 1.f, ep_m = 0.f, ep_v = 0.f;
  unsigned int ep_step = 0u;
  const bool ep_on = (a.grad_out || a.loss_out || a.fit.kind != SB_OPT_NONE) && tid < D * C::K;
  if (ep_on) {
    ep_xi = __ldcg(a.xi + tid);
    if (a.mask) ep_mk = __ldcg(a.mask + tid);
    if (a.fit.kind == SB_OPT_ADAM) { ep_m = __ldcg(a.fit.m + tid); ep_v = __ldcg(a.fit.v + tid); }
  }
  if (a.fit.kind == SB_OPT_ADAM) ep_step = __ldcg(a.fit.step);

  // ---- ordered final reduction over CTAs. The partial rows (gridDim × kRow floats, contiguous, in L2) are pulled
  // into the idle tile ring with TMA bulk copies — no registers, no load/convert serialisation — then warp w adds the
  // rows b ≡ w (mod kWarps) in row order, a lane owning the column quads {lane, lane+32, ...}, and the kWarps sub-sums
  // are added in fp64 in warp order. The association is fixed by (gridDim, kWarps) only => run-to-run deterministic.
  {
    constexpr int kQuads = C::kRow / 4;
    constexpr int kChunks = (kQuads + 31) / 32;
    const uint32_t row_bytes = (uint32_t)gridDim.x * C::kRow * (uint32_t)sizeof(float);
    float* srows = reinterpret_cast<float*>(smem_raw);
    double* sub = reinterpret_cast<double*>(smem_raw + ((row_bytes + 127u) & ~127u));   // kWarps × kRow doubles
    if (tid == 0) {
      // the rows were written through the generic proxy by other SMs and observed via the ticket: order the
      // async-proxy reads of the bulk copies after them
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_expect_tx(fin_bar, row_bytes);
      constexpr uint32_t kPiece = 32768;
      for (uint32_t off = 0; off < row_bytes; off += kPiece) {
        const uint32_t len = row_bytes - off < kPiece ? row_bytes - off : kPiece;
        tma_load_1d(smem_raw + off, reinterpret_cast<const unsigned char*>(a.partial) + off, len, fin_bar);
      }
    }
    mbar_wait(fin_bar, 0);
    stamp(a.trace, 8);
    // fp32 -> fp64 conversions and DADDs run at a fraction of the FP32 rate, and so do scalar FP32 instructions next
    // to the packed ones: the rows are added as unevaluated (hi, lo) fp32 pairs with the error-free TwoSum in packed
    // f32x2 arithmetic — exact to ~2^-48 like the fp64 sum it replaces — and only the pair totals are converted
    float2 hi[kChunks][2], lo[kChunks][2];
    static_for<0, kChunks>([&](auto c) {
      static_for<0, 2>([&](auto q) { hi[c][q] = make_float2(0.f, 0.f); lo[c][q] = make_float2(0.f, 0.f); });
    });
    const float2 neg1 = make_float2(-1.f, -1.f);
    auto sub2 = [&](float2 u, float2 w) { return __ffma2_rn(w, neg1, u); };   // u - w, one rounding per lane
    auto two_sum = [&](float2& h, float2& l, float2 x) {
      const float2 s = __fadd2_rn(h, x);
      const float2 bb = sub2(s, h);
      const float2 err = __fadd2_rn(sub2(h, sub2(s, bb)), sub2(x, bb));
      h = s;
      l = __fadd2_rn(l, err);
    };
    const float4* rows = reinterpret_cast<const float4*>(srows);
#pragma unroll 2
    for (unsigned int b = wid; b < gridDim.x; b += C::kWarps) {
      static_for<0, kChunks>([&](auto cc) {
        constexpr int c = cc;
        const int p = c * 32 + lane;
        if (p < kQuads) {
          const float4 u = rows[b * kQuads + p];
          two_sum(hi[c][0], lo[c][0], make_float2(u.x, u.y));
          two_sum(hi[c][1], lo[c][1], make_float2(u.z, u.w));
        }
      });
    }
    static_for<0, kChunks>([&](auto cc) {
      constexpr int c = cc;
      const int p = c * 32 + lane;
      if (p < kQuads) {
        double* o = sub + wid * C::kRow + 4 * p;
        o[0] = (double)hi[c][0].x + (double)lo[c][0].x; o[1] = (double)hi[c][0].y + (double)lo[c][0].y;
        o[2] = (double)hi[c][1].x + (double)lo[c][1].x; o[3] = (double)hi[c][1].y + (double)lo[c][1].y;
      }
    });
    stamp(a.trace, 9);
    __syncthreads();
    for (int e = tid; e < C::NV; e += C::kThreads) {
      double v = 0.0;
#pragma unroll
      for (int wq = 0; wq < C::kWarps; ++wq) v += sub[wq * C::kRow + e];
      fin[e] = v;
    }
  }
  if (tid == 0) { fin[C::NV] = (double)a.n; *a.ticket = 0u; }
  __syncthreads();
  stamp(a.trace, 4);

  // ---- all-reduce across GPUs inside the kernel (no NCCL launch), low-latency protocol: every value travels as one
  // 16-byte line {lo32, epoch, hi32, epoch} written with a single peer store over NVLink into slot [parity][sender] of
  // EVERY rank's symmetric buffer. A line validates itself (each 8-byte half carries the epoch), so there is no fence
  // and no separate flag: the receiver polls the lines of its OWN buffer until both epochs match and adds the senders
  // in rank order (identical bits on every rank). One NVLink one-way latency instead of store + system fence + flag.
  // Two parities suffice: a rank can run at most one step ahead, because finishing step e+1 needs every peer's lines
  // of e+1, which a peer only sends after it has read the lines of step e.
  bool peer_lost = false;
  if (a.peer.world > 1) {
    constexpr int NVX = C::NV + 1;
    static_assert(NVX <= C::kThreads, "one exchanged value per thread");
    const int world = a.peer.world, rank = a.peer.rank;
    const unsigned int epoch = *a.peer.epoch + 1u;
    const int par = (int)(epoch & 1u);
    if (tid < NVX) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(fin[tid]);
      const unsigned int lo = (unsigned int)bits, hi = (unsigned int)(bits >> 32);
      for (int r = 0; r < world; ++r) {
        uint4* dst = reinterpret_cast<uint4*>(a.peer.buf[r]) + (size_t)(par * world + rank) * NVX + tid;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(epoch), "r"(hi),
                     "r"(epoch)
                     : "memory");
      }
      double v = 0.0;
      const long long t0 = clock64();
      const long long limit = a.peer.timeout_ticks;
      bool lost = false;
      for (int r = 0; r < world; ++r) {
        const uint4* src = reinterpret_cast<const uint4*>(a.peer.buf[rank]) + (size_t)(par * world + r) * NVX + tid;
        unsigned int q0, q1, q2, q3;
        do {
          asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3)
                       : "l"(src)
                       : "memory");
        } while ((q1 != epoch || q3 != epoch) && (clock64() - t0) < limit);   // bounded: never hang the GPU on a lost peer
        // a peer that never arrived must not pass silently: the sums of this call are poisoned (NaN loss / gradient) ...
        if (q1 != epoch || q3 != epoch) { v = __longlong_as_double(0x7ff8000000000000ll); lost = true; }
        v += __longlong_as_double((long long)(((unsigned long long)q2 << 32) | (unsigned long long)q0));
      }
      fin[tid] = v;
      // ... and the failure is recorded in the rank's sticky status word (epoch of the first loss), which the host reads at
      // its sync points (FitStepper.check); the optimiser update below is skipped so that Ξ, the Adam moments and the
      // resident W stay at their last good values instead of turning NaN for good
      if (lost) atomicCAS(a.peer.status, 0u, epoch);
    }
    __syncthreads();
    if (tid == 0) *a.peer.epoch = epoch;   // after the barrier: every thread has read the old value long ago
    peer_lost = (*reinterpret_cast<volatile unsigned int*>(a.peer.status) != 0u);
  }
  const double n_total = fin[C::NV];
  stamp(a.trace, 5);

  if (a.out) {
    for (int e = tid; e < C::NV; e += C::kThreads) {
      const double v = fin[e];
      if (e == C::NV - 1) {
        if (a.write_header) { a.out[0] = v; a.out[1] = n_total; }
      } else {
        const int i = e / C::K, k = e % C::K;
        a.out[a.out_off + (a.out_transposed ? (int64_t)k * D + i : (int64_t)e)] = v;
      }
    }
  }

  // ---- closure epilogue (`train.py:663-664,680-683,689`) [+ optimiser.step() of the Adam loop, `train.py:528-530`] ----
  if (a.grad_out || a.loss_out || a.fit.kind != SB_OPT_NONE) {
    const double denom = (n_total > 0.0 ? n_total : 1.0) * D;
    // Adam bias corrections of this step (torch.optim.Adam: step counts from 1)
    float step_size = a.fit.lr, bc2_sqrt = 1.f;
    if (a.fit.kind == SB_OPT_ADAM) {
      const unsigned int t = ep_step + 1u;
      step_size = (float)((double)a.fit.lr / (1.0 - powi((double)a.fit.beta1, t)));
      bc2_sqrt = (float)sqrt(1.0 - powi((double)a.fit.beta2, t));
    }
    // linear Lie-derivative regulariser (`train.py:503-507`, intended formula) for a FIXED data set: with G = ΘᵀΘ
    // known, Σ_v Σ_n ‖J_h(z_n)(v z_n) − v h(z_n)‖² = wᵀHw is an exact quadratic form of w = vec(Ξ⊙mask); H is built
    // once per fit (symreg.quadratic_form). Here: hw = H·w by 256 threads — column quads × row groups, every load
    // of a thread in flight at once — then loss += w_sym·wᵀhw and grad += w_sym·2·hw⊙mask.
    double sym_hw = 0.0, sym_w = 0.0;
    if (a.sym_H) {
      constexpr int DK = D * C::K;
      float* wsm = reinterpret_cast<float*>(smem_raw);          // the tile ring is idle: w, then the partial products
      if (tid < DK) wsm[tid] = ep_xi * ep_mk;
      __syncthreads();
      if constexpr (DK % 4 == 0) {
        constexpr int NCQ = DK / 4, NRG = C::kThreads / NCQ, ROWS = (DK + NRG - 1) / NRG;
        float* part = wsm + ((DK + 3) & ~3);                     // [NRG][DK]
        const int rg = tid / NCQ, cq = tid % NCQ;
        if (rg < NRG) {
          const float4* Hq = reinterpret_cast<const float4*>(a.sym_H);
          float4 h[ROWS];
          static_for<0, ROWS>([&](auto rc) {
            constexpr int r = rc;
            const int j = rg + r * NRG;
            h[r] = j < DK ? __ldg(Hq + (size_t)j * NCQ + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
          });
          float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
          static_for<0, ROWS>([&](auto rc) {
            constexpr int r = rc;
            const int j = rg + r * NRG;
            const float wj = j < DK ? wsm[j] : 0.f;
            sacc.x = fmaf(h[r].x, wj, sacc.x); sacc.y = fmaf(h[r].y, wj, sacc.y);
            sacc.z = fmaf(h[r].z, wj, sacc.z); sacc.w = fmaf(h[r].w, wj, sacc.w);
          });
          reinterpret_cast<float4*>(part + rg * DK)[cq] = sacc;
        }
        __syncthreads();
        if (tid < DK) {
          for (int g = 0; g < NRG; ++g) sym_hw += (double)part[g * DK + tid];
        }
      } else {
        if (tid < DK) {   // small libraries whose d·K is not a multiple of 4: one column per thread (H is symmetric)
          for (int j = 0; j < DK; ++j) sym_hw += (double)__ldg(a.sym_H + (size_t)j * DK + tid) * (double)wsm[j];
        }
      }
      if (tid < DK) sym_w = (double)wsm[tid];
    }
    double l1 = 0.0, lsym = 0.0;
    if (ep_on) {
      const int e = tid;
      const float xi = ep_xi, mk = ep_mk;
      l1 = fabs((double)xi);
      lsym = sym_w * sym_hw;
      const double sgn = (xi > 0.f) ? 1.0 : ((xi < 0.f) ? -1.0 : 0.0);
      const float g = (float)(fin[e] * (2.0 / denom) * a.w_mse * (double)mk + a.w_l1 * sgn +
                              a.w_sym * 2.0 * sym_hw * (double)mk);
      if (a.grad_out) a.grad_out[e] = g;
      if (a.fit.kind != SB_OPT_NONE && !peer_lost) {
        float xn;
        if (a.fit.kind == SB_OPT_ADAM) {
          // the arithmetic of torch's Adam in fp32: lerp, mul+addcmul, sqrt / sqrt(bc2) + eps, addcdiv
          const float m = fmaf(1.f - a.fit.beta1, g - ep_m, ep_m);
          const float v = fmaf((1.f - a.fit.beta2) * g, g, ep_v * a.fit.beta2);
          a.fit.m[e] = m;
          a.fit.v[e] = v;
          const float den = __fdiv_rn(__fsqrt_rn(v), bc2_sqrt) + a.fit.eps;
          xn = xi - step_size * __fdiv_rn(m, den);
        } else {
          xn = xi - a.fit.lr * g;
        }
        a.fit.xi[e] = xn;
        const int i = e / C::K, k = e % C::K;
        a.w_const[i * 2 * C::K2 + k] = xn * mk;   // W of the next launch (odd-K padding stays zero)
      }
    }
    l1 = warp_sum(l1);
    lsym = warp_sum(lsym);
    __shared__ double l1w[C::kWarps], lsw[C::kWarps];
    if (lane == 0) { l1w[wid] = l1; lsw[wid] = lsym; }
    __syncthreads();
    if (tid == 0) {
      if (a.loss_out) {
        double t = 0.0, ts = 0.0;
        for (int wq = 0; wq < C::kWarps; ++wq) { t += l1w[wq]; ts += lsw[wq]; }
        *a.loss_out = (float)(a.w_mse * fin[C::NV - 1] / denom + a.w_l1 * t + a.w_sym * ts);
      }
      if (a.fit.kind == SB_OPT_ADAM && !peer_lost) *a.fit.step = ep_step + 1u;
    }
  }
  stamp(a.trace, 6);
}

// Ξ (d×K fp32) [⊙ mask] -> the packed constant slot of the fused kernels (pairs, zero padded for odd K)
__global__ void pack_w_kernel(const float* __restrict__ xi, const float* __restrict__ mask, float* __restrict__ dst,
                              int d, int K, int K2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d * K2 * 2) return;
  const int i = t / (2 * K2), k = t % (2 * K2);
  float v = 0.f;
  if (k < K) {
    v = xi[i * K + k];
    if (mask) v *= mask[i * K + k];
  }
  dst[t] = v;
}

// ---- host side ------------------------------------------------------------------------------------
std::mutex g_init_mutex;   // guards the lazily built per-device tables below (the library is re-entrant)

int tuning_variant() {
  static std::atomic<int> cached{-1};
  int v = cached.load(std::memory_order_acquire);
  if (v < 0) {
    const char* e = getenv("SB_FUSED_VARIANT");
    // A/B on B200 at N = 1e8 (one launch, median of 15): 15: 1.335 ms, 11: 1.365, 47: 1.387, 143: 1.378; at the 8-GPU
    // shard size (1.25e7 samples, whole fit step): 15: 174.9 us, 11: 180.3, 23: 177.7, 19: 183.6. Earlier rounds of the
    // same A/B (CTA barrier per tile, one prediction chain, 1024x4 ring, unroll by 2: spills) were slower and are gone,
    // and so is a packed-monomial schedule (23 mul.f32x2 + 6 FMUL instead of 52 FMUL, same bits): 1.367 ms against
    // 1.342 — a scalar FMUL holds the FMA pipe one cycle, a packed one two, so packing only saves issue slots. Reading
    // only x of the next sample ahead (3 registers; its LDS shows as the loop's largest short-scoreboard stall) made no
    // difference either: 1.3325 against 1.3300. Later (whole fit step at N = 1e8): refilling the stage consumed TWO tiles ago
    // (bit 6) 1.3008 against 1.3011, with a fourth 2048-sample stage 1.348; 4096 x 2 tiles 1.3124 — the ring is not it.
    v = e ? atoi(e) : 15;
    if (v < 0 || v > 1023) v = 15;
    cached.store(v, std::memory_order_release);
  }
  return v;
}

int small_variant() {
  static std::atomic<int> cached{-1};
  int v = cached.load(std::memory_order_acquire);
  if (v < 0) {
    const char* e = getenv("SB_FUSED_VARIANT_SMALL");
    v = e ? atoi(e) : 0;                       // 0 = per-shape default
    if (v != 5 && v != 13 && v != 37 && v != 45 && v != 21 && v != 53) v = 0;
    cached.store(v, std::memory_order_release);
  }
  return v;
}

template <int D, int P, int LEFT, int VAR, int SLOT = 0>
int launch_fused_var(FusedArgs a, void* ws, int64_t ws_bytes, cudaStream_t s, int slot = 0) {
  using C = Cfg<D, P, VAR>;
  if constexpr (LEFT == LEFT_RESIDUAL && SLOT == 0) {   // the residual kernel exists once per coefficient slot
    if (slot == 1) return launch_fused_var<D, P, LEFT, VAR, 1>(a, ws, ws_bytes, s, 1);
  }
  auto kern = fused_step_kernel<D, P, LEFT, VAR, SLOT>;
  static int grid_cached[64] = {0};  // per device (and per instantiation: this is a function template)
  int dev = 0;
  SB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return SB_ERR_INVALID; }
  std::unique_lock<std::mutex> lock(g_init_mutex);
  if (grid_cached[dev] == 0) {
    SB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    int per_sm = 0, sms = 0;
    SB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::kThreads, C::kSmemBytes));
    SB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm < 1) { set_error("fused kernel <%d,%d> does not fit on an SM", D, P); return SB_ERR_CUDA; }
    int g = per_sm * sms;
    if (g > kMaxPartialBlocks) g = kMaxPartialBlocks;
    grid_cached[dev] = g;
  }
  const int grid_max = grid_cached[dev];
  lock.unlock();
  a.n_bulk = a.n & ~(int64_t)3;
  a.n_tiles = (a.n_bulk + C::kTile - 1) / C::kTile;
  // one contiguous, equally sized range per CTA; no more CTAs than half-tiles of work
  int64_t grid = (a.n_bulk + C::kTile / 2 - 1) / (C::kTile / 2);
  if (grid > grid_max) grid = grid_max;
  // the last block stages all partial rows (+ kWarps sub-sum rows of doubles) in the tile ring
  constexpr int64_t kRowsFit = ((int64_t)C::kSmemData - 128 - (int64_t)C::kWarps * C::kRow * 8) / (C::kRow * 4);
  if (grid > kRowsFit) grid = kRowsFit;
  if (grid < 1) grid = 1;
  const int64_t need = kWsHeaderBytes + grid * C::kRow * (int64_t)sizeof(float);
  if (ws_bytes < need) {
    set_error("workspace too small: %lld < %lld bytes", (long long)ws_bytes, (long long)need);
    return SB_ERR_WORKSPACE;
  }
  a.ticket = reinterpret_cast<unsigned int*>(ws);
  a.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsHeaderBytes);
  a.trace = g_trace;
  // Programmatic dependent launch between consecutive iterations was tried (griddepcontrol.launch_dependents / .wait
  // after the first tile is in flight, cudaLaunchAttributeProgrammaticStreamSerialization): 173.11 -> 172.84 us per
  // iteration at the 8-GPU shard size, 1298.95 -> 1298.11 at N = 1e8 — the SM that hosts the predecessor's last block
  // starts last and sets the end of the successor — and the 200-iteration loss history was NO LONGER bitwise the
  // default's: the successor reads W through the constant bank while the predecessor rewrites the slot. Removed.
  kern<<<(unsigned)grid, C::kThreads, C::kSmemBytes, s>>>(a);
  SB_LAUNCH_CHECK("fused_step_kernel");
  return SB_OK;
}

template <int D, int P, int LEFT>
int launch_fused(const FusedArgs& a, void* ws, int64_t ws_bytes, cudaStream_t s, int slot = 0) {
  if constexpr (D == 3 && P == 5) {  // the headline shape carries the A/B variants
    switch (tuning_variant()) {
      case 11: return launch_fused_var<D, P, LEFT, 11>(a, ws, ws_bytes, s, slot);
      case 19: return launch_fused_var<D, P, LEFT, 19>(a, ws, ws_bytes, s, slot);
      case 23: return launch_fused_var<D, P, LEFT, 23>(a, ws, ws_bytes, s, slot);
      case 47: return launch_fused_var<D, P, LEFT, 47>(a, ws, ws_bytes, s, slot);
      case 143: return launch_fused_var<D, P, LEFT, 143>(a, ws, ws_bytes, s, slot);
      default: return launch_fused_var<D, P, LEFT, 15>(a, ws, ws_bytes, s, slot);
    }
  }
  // small libraries: empty-barrier ring + folded reduction; tile ring / prefetch per shape from the A/B on B200 at
  // N = 1e8 (median of 15, ms; SB_FUSED_VARIANT_SMALL overrides):   1024x4   2048x3   1024x4+prefetch   2048x3+prefetch
  //                                                      (2,2) K=6   0.2725   0.2626       0.2758            0.2663
  //                                                      (2,3) K=10  0.3491   0.3260       0.3237            0.2967
  //                                                      (3,3) K=20  0.5957   0.5790       0.6455            0.6731
  // 2048-sample tiles halve the per-tile barrier traffic (4 -> 8 samples per thread and tile); reading the next
  // sample ahead pays where the loop is short enough to be latency-bound (K = 10) and costs registers / issue slots
  // where the FMA pipe is the limit (K = 20). Two samples per thread and iteration (two independent chains) was also
  // tried on the 2048x3 ring: (2,3) 0.327, (3,3) 0.686 — slower (170 registers, the second expansion competes for the
  // same FMA pipe); removed. A 2048 x 2 ring (room for a second resident CTA at d = 3: 16 warps per SM instead of 8) is
  // a wash: (3,3) 0.556 against 0.560, (2,3) 0.322 against 0.290 — occupancy is not what limits these shapes; removed.
  // Round 2, later: 4096-sample tiles (16 samples per thread and tile, one resident CTA) beat both — (2,3) 0.2921 ->
  // 0.2831 with prefetch (variant 53; 0.3184 without), (3,3) 0.5624 -> 0.5329 without (variant 21; 0.5552 with), (2,2)
  // 0.2634 -> 0.2744 / 0.2667 (stays on 2048 x 3); a third 4096-sample stage at d = 2 changes nothing (0.2826 / 0.2599); two
  // prediction chains on the 4096-sample tiles are slower ((3,3) 0.5527, (2,3) 0.2954 with prefetch).
  // (3,2) K = 10: 0.3765 ms on 2048 x 3 = 0.97 of the HBM peak already; (2,2,exp) K = 8 (code 12): 0.3327 -> 0.3076 on 53.
  const int chosen = small_variant() ? small_variant()
                                     : ((D == 2 && (P == 3 || P == 12)) ? 53 : ((D == 3 && P == 3) ? 21 : 13));
  switch (chosen) {
    case 13: return launch_fused_var<D, P, LEFT, 13>(a, ws, ws_bytes, s, slot);   // 2048 x 3 ring
    case 37: return launch_fused_var<D, P, LEFT, 37>(a, ws, ws_bytes, s, slot);   // 1024 x 4 ring + next-sample prefetch
    case 45: return launch_fused_var<D, P, LEFT, 45>(a, ws, ws_bytes, s, slot);   // 2048 x 3 ring + prefetch
    case 21: return launch_fused_var<D, P, LEFT, 21>(a, ws, ws_bytes, s, slot);   // 4096 x 2 ring
    case 53: return launch_fused_var<D, P, LEFT, 53>(a, ws, ws_bytes, s, slot);   // 4096 x 2 ring + prefetch
    default: return launch_fused_var<D, P, LEFT, 5>(a, ws, ws_bytes, s, slot);
  }
}

// device (global-alias) address of constant slot `slot` on the current device
int const_slot(int slot, float** out) {
  static float* base[64] = {nullptr};
  int dev = 0;
  SB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return SB_ERR_INVALID; }
  if (slot < 0 || slot >= kWSlots) { set_error("coefficient slot %d out of range (0..%d)", slot, kWSlots - 1); return SB_ERR_INVALID; }
  std::lock_guard<std::mutex> lock(g_init_mutex);
  if (!base[dev]) {
    void* p = nullptr;
    SB_CUDA_TRY(cudaGetSymbolAddress(&p, c_w2));
    base[dev] = reinterpret_cast<float*>(p);
  }
  *out = base[dev] + (size_t)slot * kWSlotFloats;
  return SB_OK;
}

// The scratch slot (0) is shared by every stateless call on the device. Calls on ONE stream are ordered by the stream.
// A call arriving on a different stream than the previous user's must not repack the slot while that user's kernel may
// still be reading it: it waits for an event recorded behind the previous user (no-op in the common single-stream case;
// skipped while either stream is being captured into a CUDA graph — captured work replays in graph order).
struct ScratchUser { cudaStream_t stream = nullptr; cudaEvent_t ev = nullptr; bool used = false; };
ScratchUser g_scratch[64];

int guard_scratch(cudaStream_t s) {
  int dev = 0;
  SB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return SB_ERR_INVALID; }
  std::lock_guard<std::mutex> lock(g_init_mutex);
  ScratchUser& u = g_scratch[dev];
  if (u.used && u.stream != s) {
    cudaStreamCaptureStatus a = cudaStreamCaptureStatusNone, b = cudaStreamCaptureStatusNone;
    SB_CUDA_TRY(cudaStreamIsCapturing(s, &a));
    if (cudaStreamIsCapturing(u.stream, &b) != cudaSuccess) { cudaGetLastError(); b = cudaStreamCaptureStatusNone; u.used = false; }
    if (u.used && a == cudaStreamCaptureStatusNone && b == cudaStreamCaptureStatusNone) {
      if (!u.ev) SB_CUDA_TRY(cudaEventCreateWithFlags(&u.ev, cudaEventDisableTiming));
      if (cudaEventRecord(u.ev, u.stream) == cudaSuccess) SB_CUDA_TRY(cudaStreamWaitEvent(s, u.ev, 0));
      else cudaGetLastError();   // the previous user's stream no longer exists: its work has completed
    }
  }
  u.stream = s;
  u.used = true;
  return SB_OK;
}

// Ξ [⊙ mask] -> constant slot, one tiny launch on the stream (replaces a D2D cudaMemcpyToSymbolAsync + a mul)
template <int D, int P>
int upload_w(const float* xi, const float* mask, int slot, cudaStream_t s) {
  using C = Cfg<D, P>;
  static_assert(D * C::K2 * 2 <= kWSlotFloats, "a coefficient slot holds d x 2*K2 floats");
  float* base = nullptr;
  int st = const_slot(slot, &base);
  if (st != SB_OK) return st;
  if (slot == 0) { st = guard_scratch(s); if (st != SB_OK) return st; }
  const int total = D * C::K2 * 2;
  pack_w_kernel<<<(total + 255) / 256, 256, 0, s>>>(xi, mask, base, D, C::K, C::K2);
  SB_LAUNCH_CHECK("pack_w_kernel");
  return SB_OK;
}

template <int D, int P>
int run_fused(const float* x, const float* dx, int64_t n, const float* w, const float* mask, uint32_t flags,
              double* out, const ClosureOut* co, const PeerArgs* peer, void* ws, int64_t ws_bytes, cudaStream_t s,
              const FitArgs* fit) {
  using C = Cfg<D, P>;
  FusedArgs a{};
  a.x = x; a.dx = dx; a.n = n; a.out = out; a.w_mse = 1.0;
  const bool resid = flags & (SB_STEP_LOSS | SB_STEP_GRAD);
  if (resid) {
    int st = SB_OK;
    const int slot = (fit && fit->kind != SB_OPT_NONE) ? 1 : 0;
    if (!(fit && fit->w_resident)) st = upload_w<D, P>(w, mask, slot, s);
    if (st != SB_OK) return st;
    a.out_off = 2; a.out_transposed = 0; a.write_header = 1;
    if (co) {
      a.xi = w; a.mask = mask; a.w_l1 = co->w_l1; a.w_mse = co->w_mse; a.loss_out = co->loss; a.grad_out = co->grad;
    }
    if (fit && fit->kind != SB_OPT_NONE) {
      a.fit = *fit;
      a.sym_H = fit->sym_H; a.w_sym = fit->w_sym;
      st = const_slot(slot, &a.w_const);
      if (st != SB_OK) return st;
    }
    if (peer) a.peer = *peer;
    st = launch_fused<D, P, LEFT_RESIDUAL>(a, ws, ws_bytes, s, slot);
    if (st != SB_OK) return st;
    a.loss_out = nullptr; a.grad_out = nullptr; a.peer = PeerArgs{}; a.fit = FitArgs{}; a.sym_H = nullptr;
  }
  if (flags & SB_STEP_B) {
    // the ΘᵀẊ section follows the (optional) gradient and Gram sections of the packed layout
    a.out_off = 2 + ((flags & SB_STEP_GRAD) ? (int64_t)D * C::K : 0) + ((flags & SB_STEP_GRAM) ? (int64_t)C::K * C::K : 0);
    a.out_transposed = 1; a.write_header = resid ? 0 : 1;
    int st = launch_fused<D, P, LEFT_DX>(a, ws, ws_bytes, s);
    if (st != SB_OK) return st;
  }
  return SB_OK;
}

// h(x) = Θ(x)·Wᵀ per sample (`sindy.py:79-82`): Θ in registers, W pairs from the constant bank, packed FMAs.
// 8·d bytes per sample (x in, y out) against (K−1−d) + 2Kd flop: HBM/FP32 balanced for K = 56, HBM-bound below.
template <int D, int P>
__global__ void __launch_bounds__(256) forward_spec_kernel(const float* __restrict__ x, int64_t n,
                                                           float* __restrict__ y) {
  using C = Cfg<D, P, 1>;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the next sample's x is requested before the current one is expanded: ncu showed the kernel waiting on its own
  // global loads (long scoreboard 4.97 per issue, round 2) with nothing to overlap them
  float xn[D];
  if (s < n) static_for<0, D>([&](auto q) { xn[q] = __ldg(x + s * D + q); });
  for (; s < n; s += stride) {
    float xs[D], m[C::K];
    static_for<0, D>([&](auto q) { xs[q] = xn[q]; });
    const int64_t s2 = s + stride;
    if (s2 < n) static_for<0, D>([&](auto q) { xn[q] = __ldg(x + s2 * D + q); });
    expand_lib<D, P>(xs, m);
    float2 pred[D][2];
    static_for<0, D>([&](auto i) { pred[i][0] = make_float2(0.f, 0.f); pred[i][1] = make_float2(0.f, 0.f); });
    static_for<0, C::K2>([&](auto kc) {
      constexpr int kk = kc;
      const float2 m2 = make_float2(m[2 * kk], (2 * kk + 1 < C::K) ? m[2 * kk + 1] : 0.f);
      static_for<0, D>([&](auto ic) {
        constexpr int i = ic;
        pred[i][kk % 2] = __ffma2_rn(c_w2[i * C::K2 + kk], m2, pred[i][kk % 2]);
      });
    });
    static_for<0, D>([&](auto i) {
      y[s * D + i] = (pred[i][0].x + pred[i][0].y) + (pred[i][1].x + pred[i][1].y);
    });
  }
}

// The same forward with x staged through shared memory by 1-D bulk copies (the fused kernel's ring): ncu showed the
// grid-stride kernel waiting on its own global loads (long scoreboard 4.97 per issue, FMA pipe 73.6 % active) even with
// the next sample requested one iteration ahead. Persistent CTAs, tiles dealt round-robin, y written straight from
// registers (a warp's 32 samples are 32·d contiguous floats).
constexpr int kFwdTile = 2048, kFwdStages = 3, kFwdThreads = 256;   // 4096 x 2: within 1 % (0.784 / 0.441 / 0.310 ms)

template <int D, int P>
__global__ void __launch_bounds__(kFwdThreads, 2) forward_tma_kernel(const float* __restrict__ x, int64_t n,
                                                                     float* __restrict__ y) {
  using C = Cfg<D, P, 1>;
  extern __shared__ __align__(128) unsigned char fwd_smem[];
  float* tiles = reinterpret_cast<float*>(fwd_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(fwd_smem + (size_t)kFwdStages * kFwdTile * D * sizeof(float));
  uint64_t* empty = full + kFwdStages;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t n_bulk = n & ~(int64_t)3;                       // bulk copies move multiples of 16 bytes
  const int64_t n_tiles = (n_bulk + kFwdTile - 1) / kFwdTile;
  const int my_tiles = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles b, b + grid, ...
  auto tile_first = [&](int it) { return ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kFwdTile; };
  auto tile_count = [&](int it) {
    const int64_t left = n_bulk - tile_first(it);
    return (int)(left < kFwdTile ? left : kFwdTile);
  };
  auto issue = [&](int it, int stage) {
    const uint32_t bytes = (uint32_t)tile_count(it) * D * sizeof(float);
    mbar_expect_tx(&full[stage], bytes);
    tma_load_1d(tiles + (size_t)stage * kFwdTile * D, x + tile_first(it) * D, bytes, &full[stage]);
  };
  if (tid == 0) {
    for (int s = 0; s < kFwdStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kFwdThreads / 32); }
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < kFwdStages && s < my_tiles; ++s) issue(s, s);

  auto one_sample = [&](const float (&xs)[D], int64_t smp) {
    float m[C::K];
    expand_lib<D, P>(xs, m);
    float2 pred[D][2];
    static_for<0, D>([&](auto i) { pred[i][0] = make_float2(0.f, 0.f); pred[i][1] = make_float2(0.f, 0.f); });
    static_for<0, C::K2>([&](auto kc) {
      constexpr int kk = kc;
      const float2 m2 = make_float2(m[2 * kk], (2 * kk + 1 < C::K) ? m[2 * kk + 1] : 0.f);
      static_for<0, D>([&](auto ic) {
        constexpr int i = ic;
        pred[i][kk % 2] = __ffma2_rn(c_w2[i * C::K2 + kk], m2, pred[i][kk % 2]);
      });
    });
    static_for<0, D>([&](auto i) {
      y[smp * D + i] = (pred[i][0].x + pred[i][0].y) + (pred[i][1].x + pred[i][1].y);
    });
  };

  for (int it = 0; it < my_tiles; ++it) {
    const int stage = it % kFwdStages;
    if (tid == 0 && it >= 1) {                                   // refill the stage consumed one tile ago
      const int next = it + kFwdStages - 1;
      if (next < my_tiles) {
        const int ps = (it - 1) % kFwdStages;
        mbar_wait(&empty[ps], (uint32_t)((it - 1) / kFwdStages) & 1u);
        issue(next, ps);
      }
    }
    mbar_wait(&full[stage], (uint32_t)(it / kFwdStages) & 1u);
    const float* sx = tiles + (size_t)stage * kFwdTile * D;
    const int cnt = tile_count(it);
    const int64_t first = tile_first(it);
#pragma unroll 1
    for (int j = tid; j < cnt; j += kFwdThreads) {
      float xs[D];
      static_for<0, D>([&](auto q) { xs[q] = sx[j * D + q]; });
      one_sample(xs, first + j);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
  if (blockIdx.x == 0) {                                         // the n mod 4 samples no bulk copy covers
    const int64_t j = n_bulk + tid;
    if (j < n) {
      float xs[D];
      static_for<0, D>([&](auto q) { xs[q] = __ldg(x + j * D + q); });
      one_sample(xs, j);
    }
  }
}

template <int D, int P>
int run_forward(const float* x, int64_t n, const float* w, float* y, cudaStream_t s) {
  int st = upload_w<D, P>(w, nullptr, 0, s);
  if (st != SB_OK) return st;
  const char* e_ring = getenv("SB_FORWARD_RING");               // A/B switch, read per call
  const bool no_ring = e_ring && e_ring[0] == '0';
  if (n >= (int64_t)148 * 2 * kFwdTile && !no_ring && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    constexpr size_t smem = (size_t)kFwdStages * kFwdTile * D * sizeof(float) + 2 * kFwdStages * sizeof(uint64_t);
    auto kern = forward_tma_kernel<D, P>;
    static int grid_cached[64] = {0};
    int dev = 0;
    SB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return SB_ERR_INVALID; }
    {
      std::lock_guard<std::mutex> lock(g_init_mutex);
      if (grid_cached[dev] == 0) {
        SB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        SB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFwdThreads, smem));
        SB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        grid_cached[dev] = (per_sm < 1 ? 1 : per_sm) * sms;
      }
    }
    kern<<<grid_cached[dev], kFwdThreads, smem, s>>>(x, n, y);
    SB_LAUNCH_CHECK("forward_tma_kernel");
    return SB_OK;
  }
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  forward_spec_kernel<D, P><<<(unsigned)blocks, 256, 0, s>>>(x, n, y);
  SB_LAUNCH_CHECK("forward_spec_kernel");
  return SB_OK;
}

// out[i*K+k] = Σ_n g[n,i]·Θ_k(x_n) (d×K fp64): dL/dW of the forward for cotangent g (`train.py:689`) — the fused
// kernel with the cotangent in the role of dx and no prediction.
template <int D, int P>
int run_weighted_sums(const float* x, const float* g, int64_t n, double* out, void* ws, int64_t ws_bytes,
                      cudaStream_t s) {
  FusedArgs a{};
  a.x = x; a.dx = g; a.n = n; a.out = out;
  a.out_off = 0; a.out_transposed = 0; a.write_header = 0;
  return launch_fused<D, P, LEFT_DX>(a, ws, ws_bytes, s);
}

// the list of specialised libraries (polynomial only)
// (d, library code): the polynomial libraries and config 3's (2, degree 2 + exp) = code 12
#define SB_FUSED_SHAPES(X) X(2, 2) X(2, 3) X(3, 2) X(3, 3) X(3, 5) X(2, 12)

}  // namespace

void fused_set_trace(unsigned long long* p) { g_trace = p; }

bool fused_supported(const LibTab& t, uint32_t flags) {
  if (flags & SB_STEP_GRAM) return false;
  // LOSS without GRAD runs the generic residual rows (the fused kernel always accumulates the gradient)
  if ((flags & SB_STEP_LOSS) && !(flags & SB_STEP_GRAD)) return false;
#define X(D, P) if (lib_matches<D, P>(t)) return true;
  SB_FUSED_SHAPES(X)
#undef X
  return false;
}

const char* fused_variant_name(const LibTab& t, uint32_t flags) {
  (void)flags;
#define X(D, P) if (lib_matches<D, P>(t)) return "fused_tma<" #D "," #P ">";
  SB_FUSED_SHAPES(X)
#undef X
  return "generic";
}

int fused_forward(const float* x, int64_t n, const LibTab& t, const float* w, float* y, cudaStream_t s) {
#define X(D, P) if (lib_matches<D, P>(t)) return run_forward<D, P>(x, n, w, y, s);
  SB_FUSED_SHAPES(X)
#undef X
  set_error("no specialised forward for d=%d K=%d", t.d, t.K);
  return SB_ERR_UNSUPPORTED;
}

int fused_weighted_sums(const float* x, const float* g, int64_t n, const LibTab& t, double* out, void* ws,
                        int64_t ws_bytes, cudaStream_t s) {
#define X(D, P) \
  if (lib_matches<D, P>(t)) return run_weighted_sums<D, P>(x, g, n, out, ws, ws_bytes, s);
  SB_FUSED_SHAPES(X)
#undef X
  set_error("no fused kernel for d=%d K=%d", t.d, t.K);
  return SB_ERR_UNSUPPORTED;
}

int64_t fused_workspace_bytes(const LibTab& t) {
  return kWsHeaderBytes + (int64_t)kMaxPartialBlocks * ((int64_t)t.d * t.K + 4) * (int64_t)sizeof(float);
}

int fused_load_w(const LibTab& t, const float* xi, const float* mask, cudaStream_t s) {
#define X(D, P) if (lib_matches<D, P>(t)) return upload_w<D, P>(xi, mask, 1, s);
  SB_FUSED_SHAPES(X)
#undef X
  set_error("no fused kernel for d=%d K=%d", t.d, t.K);
  return SB_ERR_UNSUPPORTED;
}

int fused_train_step(const float* x, const float* dx, int64_t n, const LibTab& t, const float* w, const float* mask,
                     uint32_t flags, double* out, const ClosureOut* co, const PeerArgs* peer, void* ws,
                     int64_t ws_bytes, cudaStream_t s, const FitArgs* fit) {
#define X(D, P)                                                  \
  if (lib_matches<D, P>(t))                \
    return run_fused<D, P>(x, dx, n, w, mask, flags, out, co, peer, ws, ws_bytes, s, fit);
  SB_FUSED_SHAPES(X)
#undef X
  set_error("no fused kernel for d=%d K=%d", t.d, t.K);
  return SB_ERR_UNSUPPORTED;
}

}  // namespace sb
