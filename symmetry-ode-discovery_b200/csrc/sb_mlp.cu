// sb_mlp.cu — the frozen autoencoder's MLP on the tensor cores.
//
// Config 3's closure (`train.py:645-690` with `model_utils.py:8-67`) pushes 2·B rows through the frozen 512-wide
// encoder and decoder of `autoencoder.py:38-66` every LBFGS evaluation: value, Jacobian-vector product and the
// transpose products of both for `loss.backward()`. With the reference's call pattern those are 28 cuBLAS fp32 SIMT
// GEMMs of (2B × 512 × 512) — 68 % of the closure's GPU time on a B200 (profiles/r02_c3_closure_timing.txt). All of
// them are one operation: C = epilogue(A · Bᵀ) with B a frozen 512 × 512 matrix (W for values and tangents, Wᵀ for
// cotangents) and an epilogue that is bias + ReLU, or a ReLU mask taken from another activation tensor.
//
// `mlp_gemm_kernel`: that operation as a persistent tcgen05 kernel. fp32 faithfulness by the 3×TF32 split
// (a ≈ a_hi + a_lo, both rounded to nearest tf32, residual ≤ 2⁻²⁴|a|; A_hi·B_hi + A_hi·B_lo + A_lo·B_hi, fp32 accumulation
// in TMEM). Activations live in HBM in the PANEL FORMAT the tensor core reads: for every (128-row tile, 16-feature
// block) one contiguous 16 KB block [hi | lo], each half in the canonical K-major no-swizzle UMMA layout (8 × 16-byte
// core matrices, LBO 128 B, SBO 512 B). A stage of the operand ring is therefore TWO 1-D bulk copies (`cp.async.bulk`,
// SASS UBLKCP: 16 KB of A, 32 KB of the pre-split, pre-arranged weights) — no producer warps, no generic-proxy stores,
// no tensor maps — and the epilogue of one layer writes the next layer's operand directly (split included).
//   warp 0: bulk-copy producer (one lane), 4-stage ring        warp 1: MMA issuer (one lane), 6 MMAs of
//   128 × 256 × 8 per stage, `tcgen05.commit` to the stage's empty barrier and to the accumulator's full barrier
//   warps 2-17: epilogue (four per TMEM lane quarter, 64 columns each) — `tcgen05.ld` 16 columns at a time, bias / ReLU / mask, split,
//   16-byte stores that are 128-byte contiguous per 8 lanes.
// TMEM holds TWO 256-column accumulators of one tile: the large products (A_hi·B_hi) and the small ones, added by the
// epilogue in fp32 round-to-nearest (the tensor core's accumulator truncates; see the SPLIT comment). The last partial
// round of tiles runs as narrower pieces (runtime N), and the thin OUTPUT layer of a chain can be fused into the
// epilogue (`w_out`: per-granule partial products, summed in granule order by `mlp_sum_partials_kernel`).
// `mlp_gemm_pair_kernel`: the same tile loop on a CTA pair (`cta_group::2`), kept as a measured A/B variant.
// `mlp_in_kernel` / `mlp_out_kernel`: the thin first / last layers (2 → 512, 512 → 2 and their transposes) on CUDA
// cores, writing / reading the panel format. `mlp_pack_*`: format conversions (weights once per fit).
#include <stdlib.h>

#include "sb_common.cuh"
#include "sb_tma.cuh"

#define SB_TRY(expr) do { int _s = (expr); if (_s != SB_OK) return _s; } while (0)

namespace sb {
namespace {

constexpr int kBM = 128, kBN = 256, kBK = 16, kStages = 4;
constexpr uint32_t kABlock = kBM * kBK * 4;                    // 8 KB: hi or lo half of an activation block
constexpr uint32_t kBBlock = kBN * kBK * 4;                    // 16 KB: hi or lo half of a weight block
constexpr uint32_t kStageBytes = 2 * kABlock + 2 * kBBlock;    // 48 KB
constexpr uint32_t kLBO = 128;                                 // bytes between core matrices adjacent along K
constexpr uint32_t kSBO = (kBK / 4) * 128;                     // bytes between 8-row groups
constexpr int kEpiWarps = 16;                                  // four per TMEM lane quarter, 64 columns each
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kTmemCols = 512;
constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + 256;

enum : int { kLin = 0, kRelu = 1, kMask = 2 };
constexpr int kThinMax = 8;                                    // widest thin (first / last) layer

struct GemmArgs {
  const unsigned char* A;   // panel format, m_tiles × k_blocks blocks
  const unsigned char* W;   // packed weights, n_tiles × k_blocks blocks
  const float* bias;        // n_tiles·256 floats or null
  const unsigned char* R;   // panel format like C: mask source (kMask)
  unsigned char* C;         // panel format, m_tiles × (n_tiles·16) blocks
  int m_tiles, k_blocks, n_tiles, mode;
  const float* w_out;       // fused thin output layer (one-CTA kernel): (out_dim × n) row-major, or null
  float* partials;          // [n/64][m_pad][out_dim]: per 64-column granule the partial products of every row
  int out_dim;
  int64_t m_pad;
  int full_tiles, split, items;   // one-CTA kernel: work items = full 128×256 tiles, then the last partial round as
                                  // `split` narrower pieces per tile (128 × 256/split) so that it fills the SMs
};

struct Piece { int mt, nt, n0, bn; };
__device__ __forceinline__ Piece decode_item(int t, const GemmArgs& a) {
  if (t < a.full_tiles) return Piece{t / a.n_tiles, t % a.n_tiles, 0, kBN};
  const int u = t - a.full_tiles;
  const int tile = a.full_tiles + u / a.split, bn = kBN / a.split;
  return Piece{tile / a.n_tiles, tile % a.n_tiles, (u % a.split) * bn, bn};
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout_type [61,64) = 0 (no swizzle)
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) |
         (1ull << 46);
}
// InstrDescriptor: c_format F32 (1) [4,6); a/b format TF32 (2) [7,10),[10,13); K-major both; N>>3 [17,23); M>>4 [24,29)
constexpr uint32_t kIdesc =
    (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate,
                                         uint32_t idesc = kIdesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// round to nearest tf32 (10 explicit mantissa bits), low 13 bits cleared: what the tensor core would otherwise truncate
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r & 0xFFFFE000u);
}

// byte offset of (row, 4 features starting at k) inside one half (hi or lo) of a block with 16 features per row
__device__ __forceinline__ uint32_t blk_off(int row, int k) {
  return (uint32_t)(row >> 3) * kSBO + (uint32_t)(k >> 2) * kLBO + (uint32_t)(row & 7) * 16u;
}
__device__ __forceinline__ void store_split(unsigned char* hi_half, uint32_t half_bytes, uint32_t off, float v0, float v1,
                                            float v2, float v3) {
  // hi = rn_tf32(v), lo = rn_tf32(v − hi): |v − hi − lo| ≤ 2⁻²⁴|v| and nothing is left for the tensor core to truncate
  // (a truncating split is biased towards zero and measured 5× the error of an fp32 GEMM; this one matches it)
  const float4 hi = make_float4(tf32_rn(v0), tf32_rn(v1), tf32_rn(v2), tf32_rn(v3));
  *reinterpret_cast<float4*>(hi_half + off) = hi;
  *reinterpret_cast<float4*>(hi_half + half_bytes + off) =
      make_float4(tf32_rn(v0 - hi.x), tf32_rn(v1 - hi.y), tf32_rn(v2 - hi.z), tf32_rn(v3 - hi.w));
}

// SPLIT: the two small products (A_lo·B_hi, A_hi·B_lo) accumulate in their own 256 TMEM columns and are added to the
// main accumulator (A_hi·B_hi) by the epilogue in fp32 round-to-nearest. The tensor core's accumulator TRUNCATES on
// every accumulation, a bias that a mean over samples does not average out; a third of the accumulations on the large
// accumulator is a third of that bias (measured: 3.3e-6 -> see DESIGN.md §4.8). Costs the second accumulator buffer,
// i.e. the overlap of the epilogue with the next tile's main loop.
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) mlp_gemm_kernel(GemmArgs a) {
  constexpr uint32_t kBufs = SPLIT ? 1u : 2u;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)kStages * kStageBytes);
  uint64_t* full = bars;                    // [kStages] bulk copies landed
  uint64_t* empty = bars + kStages;         // [kStages] the stage's MMAs retired (tcgen05.commit)
  uint64_t* acc_full = bars + 2 * kStages;  // [2] accumulator complete (tcgen05.commit)
  uint64_t* acc_empty = acc_full + 2;       // [2] accumulator drained (one arrival per epilogue warp)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], kEpiWarps); mbar_init(&acc_empty[1], kEpiWarps);
    fence_mbar_init();
  }
  __syncwarp();
  if (wid == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const int KB = a.k_blocks;

  if (wid == 0) {
    // ===== bulk-copy producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < a.items; t += gridDim.x) {
        const Piece pc = decode_item(t, a);
        const unsigned char* Ap = a.A + (size_t)pc.mt * KB * (2 * kABlock);
        // rows [n0, n0 + bn) of a weight block half are contiguous: (n0/8) row groups of 512 B in
        const unsigned char* Wp = a.W + (size_t)pc.nt * KB * (2 * kBBlock) + (size_t)pc.n0 * (kBK * 4);
        const uint32_t b_bytes = (uint32_t)pc.bn * (kBK * 4);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % kStages;
          if (it >= kStages) mbar_wait(&empty[s], ((it / kStages) - 1) & 1u);
          unsigned char* st = base + (size_t)s * kStageBytes;
          mbar_expect_tx(&full[s], 2 * kABlock + 2 * b_bytes);
          tma_load_1d(st, Ap + (size_t)kb * (2 * kABlock), 2 * kABlock, &full[s]);
          if (pc.bn == kBN) {
            tma_load_1d(st + 2 * kABlock, Wp + (size_t)kb * (2 * kBBlock), 2 * kBBlock, &full[s]);
          } else {
            tma_load_1d(st + 2 * kABlock, Wp + (size_t)kb * (2 * kBBlock), b_bytes, &full[s]);
            tma_load_1d(st + 2 * kABlock + kBBlock, Wp + (size_t)kb * (2 * kBBlock) + kBBlock, b_bytes, &full[s]);
          }
        }
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      const uint64_t desc0 = umma_desc(smem_u32(base));   // later stages: the start-address field (bytes >> 4) moves
      for (int t = blockIdx.x; t < a.items; t += gridDim.x, ++lt) {
        const uint32_t idesc = (kIdesc & ~(0x3Fu << 17)) | ((uint32_t)(decode_item(t, a).bn >> 3) << 17);
        const uint32_t buf = lt % kBufs;
        if (lt >= kBufs) mbar_wait(&acc_empty[buf], ((lt / kBufs) - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_addr = tmem_d + buf * (uint32_t)kBN;
        const uint32_t s_addr = SPLIT ? tmem_d + (uint32_t)kBN : d_addr;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % kStages;
          mbar_wait(&full[s], (it / kStages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t d0 = desc0 + (uint64_t)(s * (kStageBytes >> 4));
#pragma unroll
          for (int ks = 0; ks < kBK / 8; ++ks) {
            const uint32_t adv = (uint32_t)ks * 2u * kLBO;             // 8 tf32 = two core matrices along K
            const uint64_t a_hi = d0 + (adv >> 4), a_lo = d0 + ((kABlock + adv) >> 4);
            const uint64_t b_hi = d0 + ((2 * kABlock + adv) >> 4), b_lo = d0 + ((2 * kABlock + kBBlock + adv) >> 4);
            const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
            mma_tf32(s_addr, a_lo, b_hi, acc, idesc);                   // small terms first
            mma_tf32(s_addr, a_hi, b_lo, 1u, idesc);
            mma_tf32(d_addr, a_hi, b_hi, SPLIT ? acc : 1u, idesc);
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===== epilogue: warp w reads TMEM lanes 32·(w mod 4) .. +31 =====
    const int q = wid & 3;
    const int col_lo = ((wid - 2) >> 2) * (kBN / (kEpiWarps / 4));
    const int row = 32 * q + lane;
    const uint32_t row_off = blk_off(row, 0);
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < a.items; t += gridDim.x, ++lt) {
      const Piece pc = decode_item(t, a);
      const uint32_t buf = lt % kBufs;
      mbar_wait(&acc_full[buf], (lt / kBufs) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // accumulator column c of a piece is output feature nt·256 + n0 + c
      const size_t panel = ((size_t)pc.mt * a.n_tiles * (kBN / kBK) + (size_t)pc.nt * (kBN / kBK) + (size_t)(pc.n0 / kBK)) *
                           (2 * kABlock);
      unsigned char* Cp = a.C + panel;
      const unsigned char* Rp = a.R ? a.R + panel : nullptr;
      const float* bias = a.bias ? a.bias + pc.nt * kBN + pc.n0 : nullptr;
      const int col_hi = col_lo + kBN / (kEpiWarps / 4) < pc.bn ? col_lo + kBN / (kEpiWarps / 4) : pc.bn;
      float oacc[kThinMax];
#pragma unroll
      for (int o = 0; o < kThinMax; ++o) oacc[o] = 0.f;
#pragma unroll 1
      for (int c0 = col_lo; c0 < col_hi; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + buf * (uint32_t)kBN + (uint32_t)c0, r);
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
        if constexpr (SPLIT) {
          tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)kBN + (uint32_t)c0, r);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(r[e]);
        }
        const size_t blk = (size_t)(c0 / kBK) * (2 * kABlock);
        if (bias) {
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + e4);
            v[4 * e4] += b4.x; v[4 * e4 + 1] += b4.y; v[4 * e4 + 2] += b4.z; v[4 * e4 + 3] += b4.w;
          }
        }
        if (a.mode == kRelu) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
        } else if (a.mode == kMask) {
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(Rp + blk + row_off + e4 * kLBO));
            v[4 * e4] = m4.x > 0.f ? v[4 * e4] : 0.f;
            v[4 * e4 + 1] = m4.y > 0.f ? v[4 * e4 + 1] : 0.f;
            v[4 * e4 + 2] = m4.z > 0.f ? v[4 * e4 + 2] : 0.f;
            v[4 * e4 + 3] = m4.w > 0.f ? v[4 * e4 + 3] : 0.f;
          }
        }
        if (a.C) {
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4)
            store_split(Cp + blk, kABlock, row_off + e4 * kLBO, v[4 * e4], v[4 * e4 + 1], v[4 * e4 + 2], v[4 * e4 + 3]);
        }
        if (a.w_out) {
          // fused thin output layer: per 64-column granule the row's partial products with the (out_dim × n) matrix;
          // `mlp_sum_partials_kernel` adds the granules in index order — the association is fixed by the column
          // partition alone, whatever tile or piece computed a granule
          const int ncol = pc.nt * kBN + pc.n0 + c0;
          const int n_all = a.n_tiles * kBN;
          auto dot = [&](int o) {
            const float4* wp = reinterpret_cast<const float4*>(a.w_out + (size_t)o * n_all + ncol);
            float sacc = oacc[o];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 w4 = __ldg(wp + e4);
              sacc = fmaf(v[4 * e4], w4.x, sacc); sacc = fmaf(v[4 * e4 + 1], w4.y, sacc);
              sacc = fmaf(v[4 * e4 + 2], w4.z, sacc); sacc = fmaf(v[4 * e4 + 3], w4.w, sacc);
            }
            oacc[o] = sacc;
          };
          dot(0);
          if (a.out_dim > 1) dot(1);
          if (a.out_dim > 2) { dot(2); if (a.out_dim > 3) dot(3); }
          if (a.out_dim > 4) { dot(4); if (a.out_dim > 5) dot(5); if (a.out_dim > 6) dot(6); if (a.out_dim > 7) dot(7); }
          if (((c0 + 16) & 63) == 0) {
            float* dst = a.partials + ((size_t)((ncol + 16) / 64 - 1) * a.m_pad + (size_t)pc.mt * kBM + row) * a.out_dim;
#pragma unroll
            for (int o = 0; o < kThinMax; ++o)
              if (o < a.out_dim) { dst[o] = oacc[o]; oacc[o] = 0.f; }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  __syncthreads();
  if (wid == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (`cta_group::2`): the two SMs of a TPC compute one 256 × 256 tile. Each CTA stages ITS 128 rows of A
// and ITS 128-row half of the weight block; the pair's tensor cores read both halves of B, so a CTA's shared-memory pipe
// carries 3·(4+4) KB of operand reads + 16 KB of bulk-copy writes per 8-deep step (104 B/clk) instead of 3·(4+8) + 24
// (156 B/clk of a 128 B/clk pipe: the limit ncu named for the one-CTA kernel). Rank 0 issues the MMAs; rank 1's MMA
// warp relays "my stage landed" to rank 0's barrier (plain bulk copies cannot signal a peer's mbarrier); commits are
// multicast to both CTAs; the epilogue warps of both CTAs drain their own TMEM and release rank 0's accumulator barrier.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStages2 = 6;
constexpr uint32_t kBHalf = kBBlock / 2;                              // 8 KB: 128 of the 256 weight rows, hi or lo
constexpr uint32_t kStageBytes2 = 2 * kABlock + 2 * kBHalf;           // 32 KB
constexpr size_t kSmem2 = 1024 + (size_t)kStages2 * kStageBytes2 + 256;
constexpr uint32_t kIdesc2 =
    (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// Waits use the default (CTA-scope acquire) form: what crosses the pair is consumed by the tensor core through the async
// proxy or is TMEM, never generic loads of the waiting thread. A cluster-scope acquire compiles to TRYWAIT + CCTL.IVALL —
// an L1 invalidation per stage on the MMA-issuing thread — and measured 0.150 ms per layer against 0.109 for one CTA.
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc2), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_gemm_pair_kernel(GemmArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)kStages2 * kStageBytes2);
  uint64_t* full = bars;                      // [kStages2] this CTA's copies landed (+ on rank 0: the peer's relay)
  uint64_t* empty = bars + kStages2;          // [kStages2] the stage's MMAs retired (multicast commit)
  uint64_t* acc_full = bars + 2 * kStages2;   // accumulators complete (multicast commit)
  uint64_t* acc_empty = acc_full + 1;         // rank 0: accumulators drained by the epilogue warps of BOTH CTAs
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  if (tid == 0) {
    for (int s = 0; s < kStages2; ++s) { mbar_init(&full[s], rank == 0 ? 2 : 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * kEpiWarps);
    fence_mbar_init();
  }
  __syncwarp();
  if (wid == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                          // both CTAs' barriers exist before any remote arrive / multicast commit
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const int m_pairs = (a.m_tiles + 1) >> 1;
  const int n_tiles_total = m_pairs * a.n_tiles;
  const int KB = a.k_blocks;

  if (wid == 0) {
    // ===== bulk-copy producer (both CTAs: own A rows, own half of the weight rows) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = pair; t < n_tiles_total; t += n_pairs) {
        const int mp = t / a.n_tiles, nt = t % a.n_tiles;
        int mt = 2 * mp + (int)rank;
        if (mt >= a.m_tiles) mt = a.m_tiles - 1;                       // odd tile count: the idle half re-reads valid rows
        const unsigned char* Ap = a.A + (size_t)mt * KB * (2 * kABlock);
        const unsigned char* Wp = a.W + (size_t)nt * KB * (2 * kBBlock) + (size_t)rank * kBHalf;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % kStages2;
          if (it >= kStages2) mbar_wait(&empty[s], ((it / kStages2) - 1) & 1u);
          unsigned char* st = base + (size_t)s * kStageBytes2;
          mbar_expect_tx(&full[s], kStageBytes2);
          tma_load_1d(st, Ap + (size_t)kb * (2 * kABlock), 2 * kABlock, &full[s]);
          tma_load_1d(st + 2 * kABlock, Wp + (size_t)kb * (2 * kBBlock), kBHalf, &full[s]);
          tma_load_1d(st + 2 * kABlock + kBHalf, Wp + (size_t)kb * (2 * kBBlock) + kBBlock, kBHalf, &full[s]);
        }
      }
    }
  } else if (wid == 1) {
    if (lane == 0 && rank == 1) {
      // ===== relay: tell rank 0 that this CTA's half of the stage has landed =====
      uint32_t it = 0;
      for (int t = pair; t < n_tiles_total; t += n_pairs)
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % kStages2;
          mbar_wait(&full[s], (it / kStages2) & 1u);
          mbar_arrive_remote(&full[s], 0);
        }
    } else if (lane == 0) {
      // ===== MMA issuer (rank 0) =====
      uint32_t it = 0, lt = 0;
      const uint64_t desc0 = umma_desc(smem_u32(base));
      const uint32_t d_addr = tmem_d, s_addr = tmem_d + (uint32_t)kBN;
      for (int t = pair; t < n_tiles_total; t += n_pairs, ++lt) {
        if (lt >= 1) mbar_wait(acc_empty, (lt - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % kStages2;
          mbar_wait(&full[s], (it / kStages2) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t d0 = desc0 + (uint64_t)(s * (kStageBytes2 >> 4));
#pragma unroll
          for (int ks = 0; ks < kBK / 8; ++ks) {
            const uint32_t adv = (uint32_t)ks * 2u * kLBO;
            const uint64_t a_hi = d0 + (adv >> 4), a_lo = d0 + ((kABlock + adv) >> 4);
            const uint64_t b_hi = d0 + ((2 * kABlock + adv) >> 4), b_lo = d0 + ((2 * kABlock + kBHalf + adv) >> 4);
            const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
            mma_tf32_pair(s_addr, a_lo, b_hi, acc);
            mma_tf32_pair(s_addr, a_hi, b_lo, 1u);
            mma_tf32_pair(d_addr, a_hi, b_hi, acc);
          }
          mma_commit_pair(&empty[s]);
        }
        mma_commit_pair(acc_full);
      }
    }
  } else {
    // ===== epilogue (both CTAs, own 128 rows) =====
    const int q = wid & 3;
    const int col_lo = ((wid - 2) >> 2) * (kBN / (kEpiWarps / 4));
    const int row = 32 * q + lane;
    const uint32_t row_off = blk_off(row, 0);
    uint32_t lt = 0;
    for (int t = pair; t < n_tiles_total; t += n_pairs, ++lt) {
      const int mp = t / a.n_tiles, nt = t % a.n_tiles;
      const int mt = 2 * mp + (int)rank;
      const bool live = mt < a.m_tiles;
      mbar_wait(acc_full, lt & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const size_t panel = ((size_t)(live ? mt : 0) * a.n_tiles * (kBN / kBK) + (size_t)nt * (kBN / kBK)) * (2 * kABlock);
      unsigned char* Cp = a.C + panel;
      const unsigned char* Rp = a.R ? a.R + panel : nullptr;
      const float* bias = a.bias ? a.bias + nt * kBN : nullptr;
      if (live) {
#pragma unroll 1
        for (int c0 = col_lo; c0 < col_lo + kBN / (kEpiWarps / 4); c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, r);
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
          tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)kBN + (uint32_t)c0, r);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(r[e]);
          const size_t blk = (size_t)(c0 / kBK) * (2 * kABlock);
          if (bias) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + e4);
              v[4 * e4] += b4.x; v[4 * e4 + 1] += b4.y; v[4 * e4 + 2] += b4.z; v[4 * e4 + 3] += b4.w;
            }
          }
          if (a.mode == kRelu) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          } else if (a.mode == kMask) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 m4 = __ldg(reinterpret_cast<const float4*>(Rp + blk + row_off + e4 * kLBO));
              v[4 * e4] = m4.x > 0.f ? v[4 * e4] : 0.f;
              v[4 * e4 + 1] = m4.y > 0.f ? v[4 * e4 + 1] : 0.f;
              v[4 * e4 + 2] = m4.z > 0.f ? v[4 * e4 + 2] : 0.f;
              v[4 * e4 + 3] = m4.w > 0.f ? v[4 * e4 + 3] : 0.f;
            }
          }
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4)
            store_split(Cp + blk, kABlock, row_off + e4 * kLBO, v[4 * e4], v[4 * e4 + 1], v[4 * e4 + 2], v[4 * e4 + 3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(acc_empty);
        else mbar_arrive_remote(acc_empty, 0);
      }
    }
  }
  __syncthreads();
  cluster_sync_all();                          // no CTA leaves while its partner may still read its shared memory
  if (wid == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

// ---- thin first layer: out[m][f] = epilogue(Σ_i x[m][i]·w[f][i] (+ bias[f])), i < in_dim ≤ 8, panel-format output ----
struct InArgs {
  const float* x; int64_t m; int in_dim;
  const float* w; const float* bias; const unsigned char* R; unsigned char* C;
  int F, mode;
};

__global__ void __launch_bounds__(128) mlp_in_kernel(InArgs a) {
  __shared__ float ws[128 * kThinMax];
  __shared__ float bs[128];
  const int row = threadIdx.x, mt = blockIdx.x, f0 = blockIdx.y * 128;
  for (int i = row; i < 128 * a.in_dim; i += 128) ws[i] = a.w[(size_t)f0 * a.in_dim + i];
  bs[row] = a.bias ? a.bias[f0 + row] : 0.f;
  __syncthreads();
  const int64_t m = (int64_t)mt * 128 + row;
  float xr[kThinMax];
#pragma unroll
  for (int i = 0; i < kThinMax; ++i) xr[i] = (i < a.in_dim && m < a.m) ? a.x[m * a.in_dim + i] : 0.f;
  const size_t panel = ((size_t)mt * (a.F / kBK) + (size_t)(f0 / kBK)) * (2 * kABlock);
  const uint32_t row_off = blk_off(row, 0);
#pragma unroll 1
  for (int ob = 0; ob < 128 / kBK; ++ob) {
    const size_t blk = panel + (size_t)ob * (2 * kABlock);
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int fl = ob * kBK + kq * 4 + e;
        float acc = bs[fl];
#pragma unroll
        for (int i = 0; i < kThinMax; ++i)
          if (i < a.in_dim) acc = fmaf(xr[i], ws[fl * a.in_dim + i], acc);
        v[e] = acc;
      }
      if (a.mode == kRelu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
      } else if (a.mode == kMask) {
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.R + blk + row_off + kq * kLBO));
        v[0] = m4.x > 0.f ? v[0] : 0.f; v[1] = m4.y > 0.f ? v[1] : 0.f;
        v[2] = m4.z > 0.f ? v[2] : 0.f; v[3] = m4.w > 0.f ? v[3] : 0.f;
      }
      store_split(a.C + blk, kABlock, row_off + kq * kLBO, v[0], v[1], v[2], v[3]);
    }
  }
}

// ---- thin last layer: y[m][o] = Σ_f A[m][f]·w[o][f] (+ bias[o]), o < out_dim ≤ 8, panel-format input ----
struct OutArgs {
  const unsigned char* A; int64_t m; int F;
  const float* w; const float* bias; int out_dim; float* y;
};

__global__ void __launch_bounds__(512) mlp_out_kernel(OutArgs a) {
  extern __shared__ float out_smem[];
  float* ws = out_smem;                         // out_dim × F
  float* red = out_smem + a.out_dim * a.F;      // 4 × 128 × kThinMax
  const int row = threadIdx.x, part = threadIdx.y, mt = blockIdx.x;
  for (int i = part * 128 + row; i < a.out_dim * a.F; i += 512) ws[i] = a.w[i];
  __syncthreads();
  float acc[kThinMax];
#pragma unroll
  for (int o = 0; o < kThinMax; ++o) acc[o] = 0.f;
  const unsigned char* panel = a.A + (size_t)mt * (a.F / kBK) * (2 * kABlock);
  const uint32_t row_off = blk_off(row, 0);
  for (int kb = part; kb < a.F / kBK; kb += 4) {
    const unsigned char* blk = panel + (size_t)kb * (2 * kABlock);
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const float4 hi = __ldg(reinterpret_cast<const float4*>(blk + row_off + kq * kLBO));
      const float4 lo = __ldg(reinterpret_cast<const float4*>(blk + kABlock + row_off + kq * kLBO));
      const float h[4] = {hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w};
      const int f = kb * kBK + kq * 4;
#pragma unroll
      for (int o = 0; o < kThinMax; ++o)
        if (o < a.out_dim) {
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[o] = fmaf(h[e], ws[o * a.F + f + e], acc[o]);
        }
    }
  }
#pragma unroll
  for (int o = 0; o < kThinMax; ++o) red[(part * 128 + row) * kThinMax + o] = acc[o];
  __syncthreads();
  const int64_t m = (int64_t)mt * 128 + row;
  if (part == 0 && m < a.m) {
    for (int o = 0; o < a.out_dim; ++o) {
      float s = a.bias ? a.bias[o] : 0.f;
      for (int p = 0; p < 4; ++p) s += red[(p * 128 + row) * kThinMax + o];
      a.y[m * a.out_dim + o] = s;
    }
  }
}

// ---- format conversions ----
// weights: logical B[n][k] (n < N output features of the product, k < K reduction index) from w: B[n][k] = w[n·K + k],
// or, transposed, B[n][k] = w[k·N + n]
__global__ void mlp_pack_w_kernel(const float* w, int N, int K, int transpose, unsigned char* P) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (n, k-quad)
  const int kq_total = K / 4;
  if (idx >= (int64_t)N * kq_total) return;
  const int n = (int)(idx / kq_total), k = (int)(idx % kq_total) * 4;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = transpose ? w[(size_t)(k + e) * N + n] : w[(size_t)n * K + k + e];
  const int nt = n / kBN, nl = n % kBN, kb = k / kBK, kl = k % kBK;
  unsigned char* blk = P + ((size_t)nt * (K / kBK) + kb) * (2 * kBBlock);
  store_split(blk, kBBlock, blk_off(nl, kl), v[0], v[1], v[2], v[3]);
}

__global__ void mlp_pack_rows_kernel(const float* x, int64_t m, int F, unsigned char* P, int64_t m_pad) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (k-quad, row): rows fastest -> coalesced stores
  const int kq_total = F / 4;
  if (idx >= m_pad * kq_total) return;
  const int64_t mt = idx / ((int64_t)128 * kq_total);
  const int rem = (int)(idx % ((int64_t)128 * kq_total));
  const int kqi = rem / 128, row = rem % 128;
  const int64_t r = mt * 128 + row;
  const int k = kqi * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < m) v = *reinterpret_cast<const float4*>(x + r * F + k);
  unsigned char* blk = P + ((size_t)mt * (F / kBK) + k / kBK) * (2 * kABlock);
  store_split(blk, kABlock, blk_off(row, k % kBK), v.x, v.y, v.z, v.w);
}

__global__ void mlp_unpack_rows_kernel(const unsigned char* P, int64_t m, int F, float* x) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int kq_total = F / 4;
  const int64_t m_pad = (m + 127) / 128 * 128;
  if (idx >= m_pad * kq_total) return;
  const int64_t mt = idx / ((int64_t)128 * kq_total);
  const int rem = (int)(idx % ((int64_t)128 * kq_total));
  const int kqi = rem / 128, row = rem % 128;
  const int64_t r = mt * 128 + row;
  if (r >= m) return;
  const int k = kqi * 4;
  const unsigned char* blk = P + ((size_t)mt * (F / kBK) + k / kBK) * (2 * kABlock);
  const float4 hi = *reinterpret_cast<const float4*>(blk + blk_off(row, k % kBK));
  const float4 lo = *reinterpret_cast<const float4*>(blk + kABlock + blk_off(row, k % kBK));
  *reinterpret_cast<float4*>(x + r * F + k) = make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
}

// y[m][o] = (bias[o]) + Σ_g partials[g][m][o], granules in index order
__global__ void mlp_sum_partials_kernel(const float* partials, int64_t m, int64_t m_pad, int n_granules, int out_dim,
                                        const float* bias, float* y) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * out_dim) return;
  const int64_t row = idx / out_dim;
  const int o = (int)(idx % out_dim);
  float s = bias ? bias[o] : 0.f;
  for (int g = 0; g < n_granules; ++g) s += partials[((size_t)g * m_pad + row) * out_dim + o];
  y[idx] = s;
}

int check_wide(int f, const char* what) {
  if (f <= 0 || f % kBN != 0 || f > 2048) {
    set_error("sb_mlp: %s = %d must be a positive multiple of %d (<= 2048)", what, f, kBN);
    return SB_ERR_UNSUPPORTED;
  }
  return SB_OK;
}
int check_thin(int d, const char* what) {
  if (d <= 0 || d > kThinMax) {
    set_error("sb_mlp: %s = %d must be in [1, %d]", what, d, kThinMax);
    return SB_ERR_UNSUPPORTED;
  }
  return SB_OK;
}

}  // namespace

int64_t mlp_panel_bytes(int64_t m, int f) { return (m + 127) / 128 * 128 * (int64_t)f * 8; }

int mlp_pack_weights(const float* w, int n, int k, int transpose, void* packed, cudaStream_t s) {
  SB_TRY(check_wide(n, "n")); SB_TRY(check_wide(k, "k"));
  const int64_t units = (int64_t)n * (k / 4);
  mlp_pack_w_kernel<<<(unsigned)((units + 255) / 256), 256, 0, s>>>(w, n, k, transpose, (unsigned char*)packed);
  SB_LAUNCH_CHECK("mlp_pack_w_kernel");
  return SB_OK;
}

int mlp_pack_rows(const float* x, int64_t m, int f, void* packed, cudaStream_t s) {
  SB_TRY(check_wide(f, "f"));
  if (m <= 0) return SB_OK;
  const int64_t m_pad = (m + 127) / 128 * 128;
  const int64_t units = m_pad * (f / 4);
  mlp_pack_rows_kernel<<<(unsigned)((units + 255) / 256), 256, 0, s>>>(x, m, f, (unsigned char*)packed, m_pad);
  SB_LAUNCH_CHECK("mlp_pack_rows_kernel");
  return SB_OK;
}

int mlp_unpack_rows(const void* packed, int64_t m, int f, float* x, cudaStream_t s) {
  SB_TRY(check_wide(f, "f"));
  if (m <= 0) return SB_OK;
  const int64_t units = (m + 127) / 128 * 128 * (f / 4);
  mlp_unpack_rows_kernel<<<(unsigned)((units + 255) / 256), 256, 0, s>>>((const unsigned char*)packed, m, f, x);
  SB_LAUNCH_CHECK("mlp_unpack_rows_kernel");
  return SB_OK;
}

int64_t mlp_partials_bytes(int64_t m, int n, int out_dim) {
  return (int64_t)(n / 64) * ((m + 127) / 128 * 128) * out_dim * (int64_t)sizeof(float);
}

int mlp_gemm(const void* a_panel, int64_t m, int k, const void* w_packed, int n, const float* bias,
             const void* mask_panel, int mode, void* c_panel, cudaStream_t s) {
  return mlp_gemm_out(a_panel, m, k, w_packed, n, bias, mask_panel, mode, c_panel, nullptr, nullptr, 0, nullptr, nullptr, s);
}

int mlp_gemm_out(const void* a_panel, int64_t m, int k, const void* w_packed, int n, const float* bias,
                 const void* mask_panel, int mode, void* c_panel, const float* w_out, const float* bias_out,
                 int out_dim, void* partials, float* y, cudaStream_t s) {
  SB_TRY(check_wide(n, "n")); SB_TRY(check_wide(k, "k"));
  if (w_out) {
    SB_TRY(check_thin(out_dim, "out_dim"));
    if (!partials || !y) { set_error("sb_mlp_gemm_out: partials / y missing"); return SB_ERR_INVALID; }
  } else if (!c_panel) {
    set_error("sb_mlp_gemm: no output requested"); return SB_ERR_INVALID;
  }
  if (mode < kLin || mode > kMask) { set_error("sb_mlp_gemm: mode %d", mode); return SB_ERR_INVALID; }
  if (mode == kMask && !mask_panel) { set_error("sb_mlp_gemm: mask mode without a mask panel"); return SB_ERR_INVALID; }
  if (m <= 0) return SB_OK;
  if ((m + 127) / 128 * (n / kBN) > (int64_t)1 << 30) { set_error("sb_mlp_gemm: too many rows"); return SB_ERR_INVALID; }
  static bool attr_set[64] = {false};
  static int sm_count[64] = {0};
  const char* es = getenv("SB_MLP_SPLIT_ACC");
  const bool split = !(es && es[0] == '0');
  const char* ep = getenv("SB_MLP_PAIR");          // A/B switches, read per call (tests flip them)
  const bool pairs = ep && ep[0] == '1';
  const char* en = getenv("SB_MLP_NARROW");
  const bool no_narrow = en && en[0] == '0';
  int dev = 0;
  SB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("sb_mlp_gemm: device index %d", dev); return SB_ERR_INVALID; }
  if (!attr_set[dev]) {
    SB_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    SB_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    SB_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2));
    SB_CUDA_TRY(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set[dev] = true;
  }
  GemmArgs a;
  a.A = (const unsigned char*)a_panel; a.W = (const unsigned char*)w_packed; a.bias = bias;
  a.R = (const unsigned char*)mask_panel; a.C = (unsigned char*)c_panel;
  a.m_tiles = (int)((m + 127) / 128); a.k_blocks = k / kBK; a.n_tiles = n / kBN; a.mode = mode;
  a.w_out = w_out; a.partials = (float*)partials; a.out_dim = out_dim; a.m_pad = (int64_t)a.m_tiles * kBM;
  if (pairs && !w_out) {
    const int tiles2 = ((a.m_tiles + 1) / 2) * a.n_tiles;
    const int max_pairs = sm_count[dev] / 2;
    const int grid2 = 2 * (tiles2 < max_pairs ? tiles2 : max_pairs);
    mlp_gemm_pair_kernel<<<grid2, kThreads, kSmem2, s>>>(a);
    SB_LAUNCH_CHECK("mlp_gemm_pair_kernel");
    return SB_OK;
  }
  const int tiles = a.m_tiles * a.n_tiles;
  // the last partial round (626 tiles on 148 SMs = 4 rounds + 34 tiles) as narrower pieces that fill the machine
  const int sms = sm_count[dev];
  const int rem = tiles % sms;
  a.full_tiles = tiles - rem;
  a.split = 1;
  if (rem > 0 && !no_narrow) a.split = sms / rem >= 4 ? 4 : sms / rem >= 2 ? 2 : 1;
  a.items = a.full_tiles + rem * a.split;
  const int grid = a.items < sms ? a.items : sms;
  if (split) mlp_gemm_kernel<true><<<grid, kThreads, kSmem, s>>>(a);
  else mlp_gemm_kernel<false><<<grid, kThreads, kSmem, s>>>(a);
  SB_LAUNCH_CHECK("mlp_gemm_kernel");
  if (w_out) {
    const int64_t units = m * out_dim;
    mlp_sum_partials_kernel<<<(unsigned)((units + 255) / 256), 256, 0, s>>>((const float*)partials, m, a.m_pad, n / 64,
                                                                          out_dim, bias_out, y);
    SB_LAUNCH_CHECK("mlp_sum_partials_kernel");
  }
  return SB_OK;
}

int mlp_thin_in(const float* x, int64_t m, int in_dim, const float* w, const float* bias, const void* mask_panel,
                int f, int mode, void* c_panel, cudaStream_t s) {
  SB_TRY(check_wide(f, "f")); SB_TRY(check_thin(in_dim, "in_dim"));
  if (mode < kLin || mode > kMask) { set_error("sb_mlp_thin_in: mode %d", mode); return SB_ERR_INVALID; }
  if (mode == kMask && !mask_panel) { set_error("sb_mlp_thin_in: mask mode without a mask panel"); return SB_ERR_INVALID; }
  if (m <= 0) return SB_OK;
  InArgs a{x, m, in_dim, w, bias, (const unsigned char*)mask_panel, (unsigned char*)c_panel, f, mode};
  dim3 grid((unsigned)((m + 127) / 128), (unsigned)(f / 128));
  mlp_in_kernel<<<grid, 128, 0, s>>>(a);
  SB_LAUNCH_CHECK("mlp_in_kernel");
  return SB_OK;
}

int mlp_thin_out(const void* a_panel, int64_t m, int f, const float* w, const float* bias, int out_dim, float* y,
                 cudaStream_t s) {
  SB_TRY(check_wide(f, "f")); SB_TRY(check_thin(out_dim, "out_dim"));
  if (m <= 0) return SB_OK;
  OutArgs a{(const unsigned char*)a_panel, m, f, w, bias, out_dim, y};
  const size_t smem = ((size_t)out_dim * f + 4 * 128 * kThinMax) * sizeof(float);
  static bool attr_set[64] = {false};
  int dev = 0;
  SB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    SB_CUDA_TRY(cudaFuncSetAttribute(mlp_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set[dev] = true;
  }
  mlp_out_kernel<<<(unsigned)((m + 127) / 128), dim3(128, 4), smem, s>>>(a);
  SB_LAUNCH_CHECK("mlp_out_kernel");
  return SB_OK;
}

}  // namespace sb
