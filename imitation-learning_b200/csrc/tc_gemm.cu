// tcgen05 engine for the dense H x H layers of the replica-batched MLPs (sm_100a only).
//
//   C[g] (M x N) = A[g] (M x K) * B[g] (K x N), fp32 in / fp32 out, tensor-core arithmetic:
//     IL_GEMM_TF32X3 : each operand is split x = hi + lo (hi = rna_tf32(x), lo = rna_tf32(x - hi)) and the product
//                      is hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM — fp32-level accuracy (error ~2^-21
//                      per product, the dropped lo*lo term) at 3 MMAs per product ("3xTF32").
//     IL_GEMM_TF32   : hi*hi only (10-bit mantissa operands).
//
// Structure (one persistent CTA per SM, 13 warps):
//   warps 5..12  producers : global (fp32, either operand layout) -> registers -> hi/lo split -> shared memory in
//                            the canonical K-major SWIZZLE_128B UMMA layout (MN-major sources are transposed in
//                            registers with quad shuffles), fence.proxy.async, mbarrier arrive.     [2 stages x 96 KB]
//   warp  4      MMA issuer: one elected thread issues tcgen05.mma.kind::tf32 (M128 x N256 x K8) on shared-memory
//                            descriptors, accumulating into TMEM; tcgen05.commit releases stages / publishes tiles.
//   warps 0..3   epilogue  : tcgen05.ld (32 lanes x 32 columns) -> bias / activation / activation-derivative mask
//                            -> transposed through shared memory -> coalesced 128-bit global stores.
//   TMEM holds two 128 x 256 fp32 accumulators (512 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// A TMA path is not used because every operand needs the hi/lo split (a CUDA-core pass over the tile) anyway.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 32;         // tile: 128 x 256 outputs, 32-float (128 B) k-blocks
constexpr int STAGES = 2;
constexpr int N_PRODUCER_WARPS = 8, N_EPI_WARPS = 4;
constexpr int THREADS = (N_EPI_WARPS + 1 + N_PRODUCER_WARPS) * 32;  // 416
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;          // 16 KB, 32 KB (per hi or lo copy)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;               // 96 KB
constexpr int EPI_LD = 33;                                           // padded row of the epilogue staging tile
constexpr int EPI_BYTES = N_EPI_WARPS * 32 * EPI_LD * 4;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + 256;
constexpr uint32_t TMEM_COLS = 512;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO(1) | SBO(1024 B)
// | version 1 | layout SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major), in 16 B units
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

struct TcParams {
  GemmArgs g;
  int tiles_m;      // M / BM
  int n_tiles;      // G * tiles_m
  int split;        // 1: 3xTF32, 0: single TF32
};

__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 v, bool split) {
  uint4 h;
  h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
  *reinterpret_cast<uint4*>(hi_tile + off) = h;
  if (split) {
    uint4 l;
    l.x = to_tf32(v.x - __uint_as_float(h.x)); l.y = to_tf32(v.y - __uint_as_float(h.y));
    l.z = to_tf32(v.z - __uint_as_float(h.z)); l.w = to_tf32(v.w - __uint_as_float(h.w));
    *reinterpret_cast<uint4*>(lo_tile + off) = l;
  }
}

// Fills one operand tile (ROWS x 32 k) of the current stage. kmajor: element (r, k) at src[r * ld + k]; else src[k * ld + r].
template <int ROWS>
__device__ __forceinline__ void produce_tile(const float* __restrict__ src, int ld, bool kmajor, int row0, int k0, uint8_t* hi_tile, uint8_t* lo_tile, bool split, int ptid) {
  constexpr int PER_THREAD = ROWS * 8 / (N_PRODUCER_WARPS * 32);  // float4 per thread: 4 (A) or 8 (B)
  float4 v[PER_THREAD];
  if (kmajor) {
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
      const int i = ptid + j * (N_PRODUCER_WARPS * 32);
      const int r = i >> 3, c = i & 7;
      v[j] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(row0 + r) * ld + k0 + c * 4));
    }
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
      const int i = ptid + j * (N_PRODUCER_WARPS * 32);
      store_split(hi_tile, lo_tile, sw128(i >> 3, i & 7), v[j], split);
    }
  } else {
    const int pw = ptid >> 5, lane = ptid & 31, q = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
      const int it = pw + j * N_PRODUCER_WARPS;  // warp-iteration: 4 k x 32 rows
      const int kq = it & 7, rb = it >> 3;
      v[j] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(k0 + kq * 4 + t) * ld + row0 + rb * 32 + q * 4));
    }
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
      const int it = pw + j * N_PRODUCER_WARPS;
      const int kq = it & 7, rb = it >> 3;
      // 4x4 transpose inside the quad: lane t ends with (row r0 + t, k0..k0+3)
      const float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      float o[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int give = t ^ s;  // element index this lane supplies in round s
        const float mine = give == 0 ? e[0] : (give == 1 ? e[1] : (give == 2 ? e[2] : e[3]));
        const float got = __shfl_xor_sync(0xffffffffu, mine, s, 4);
        if ((t ^ s) == 0) o[0] = got;
        if ((t ^ s) == 1) o[1] = got;
        if ((t ^ s) == 2) o[2] = got;
        if ((t ^ s) == 3) o[3] = got;
      }
      store_split(hi_tile, lo_tile, sw128(rb * 32 + q * 4 + t, kq), make_float4(o[0], o[1], o[2], o[3]), split);
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  float* epi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  // bars: full[2], empty[2], tmem_full[2], tmem_empty[2], then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (4 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (6 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmArgs& g = p.g;
  const int nkb = g.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), N_PRODUCER_WARPS); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), N_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == N_EPI_WARPS) {  // TMEM allocation by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp > N_EPI_WARPS) {
    // ================= producers =================
    const int ptid = threadIdx.x - (N_EPI_WARPS + 1) * 32;
    uint32_t kb_global = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int grp = tile / p.tiles_m, m0 = (tile % p.tiles_m) * BM;
      const float* A = g.A + (int64_t)(grp / g.a_gdiv) * g.a_gs;
      const float* B = g.B + (int64_t)(grp / g.b_gdiv) * g.b_gs;
      for (int kb = 0; kb < nkb; ++kb, ++kb_global) {
        const int s = kb_global % STAGES;
        const uint32_t ph = (kb_global / STAGES) & 1;
        if (lane == 0) mbar_wait(empty_bar(s), ph ^ 1);
        __syncwarp();
        uint8_t* st = stage_base + s * STAGE_BYTES;
        produce_tile<BM>(A, g.lda, g.a_kmajor != 0, m0, kb * BK, st, st + A_BYTES, p.split != 0, ptid);
        produce_tile<BN>(B, g.ldb, g.b_kmajor != 0, 0, kb * BK, st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, p.split != 0, ptid);
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(s));
      }
    }
  } else if (warp == N_EPI_WARPS) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      // InstrDescriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major A/B, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t kb_global = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb, ++kb_global) {
          const int s = kb_global % STAGES;
          mbar_wait(full_bar(s), (kb_global / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(stage_base + s * STAGE_BYTES), a_lo = a_hi + A_BYTES, b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t ko = kk * 32;  // 8 tf32 = 32 bytes along K inside the 128 B swizzled row
            const uint32_t first = (kb == 0 && kk == 0) ? 0u : 1u;
            if (p.split) {
              tc_mma_tf32(d_tmem, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, first);
              tc_mma_tf32(d_tmem, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
              tc_mma_tf32(d_tmem, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, 1u);
            } else {
              tc_mma_tf32(d_tmem, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, first);
            }
          }
          tc_commit(empty_bar(s));  // frees the stage once these MMAs have read it (implicit before_thread_sync fence)
        }
        tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 0..3 <-> TMEM lanes 32w .. 32w+31) =================
    float* stg = epi + warp * 32 * EPI_LD;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int grp = tile / p.tiles_m, m0 = (tile % p.tiles_m) * BM;
      mbar_wait(tfull_bar(acc), (it >> 1) & 1);
      tc_fence_after();
      float* C = g.C + (int64_t)grp * g.c_gs;
      const float* bias = g.bias ? g.bias + (int64_t)grp * g.bias_gs : nullptr;
      const float* mask = g.mask ? g.mask + (int64_t)grp * g.mask_gs : nullptr;
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN + cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
            "%26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
              "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // lane = row (32 rows of this warp), registers = 32 consecutive columns -> staging tile [row][col]
#pragma unroll
        for (int c = 0; c < 32; ++c) stg[lane * EPI_LD + c] = __uint_as_float(r[c]);
        __syncwarp();
        // coalesced write-out: 8 lanes cover one 128-byte row segment, 4 rows per pass
        const int cq = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
          const int rl = pass * 4 + rsub;
          const int m = m0 + warp * 32 + rl, n = cb * 32 + cq;
          float4 v = make_float4(stg[rl * EPI_LD + cq], stg[rl * EPI_LD + cq + 1], stg[rl * EPI_LD + cq + 2], stg[rl * EPI_LD + cq + 3]);
          if (bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          }
          if (g.act >= 0) { v.x = act_apply(v.x, g.act); v.y = act_apply(v.y, g.act); v.z = act_apply(v.z, g.act); v.w = act_apply(v.w, g.act); }
          if (mask) {
            const float4 mv = __ldg(reinterpret_cast<const float4*>(mask + (int64_t)m * g.ldmask + n));
            v.x *= act_grad_from_output(mv.x, g.mask_act); v.y *= act_grad_from_output(mv.y, g.mask_act);
            v.z *= act_grad_from_output(mv.z, g.mask_act); v.w *= act_grad_from_output(mv.w, g.mask_act);
          }
          *reinterpret_cast<float4*>(C + (int64_t)m * g.ldc + n) = v;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));  // accumulator drained
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == N_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// bias gradient for the dW products routed to the tensor-core engine: out[g, n] = sum_b dY[g, b, n]
__global__ void colsum_kernel(const float* __restrict__ A, int64_t a_gs, int a_gdiv, int lda, int K, int M, float* __restrict__ out, int64_t out_gs) {
  const int g = blockIdx.y, m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* a = A + (int64_t)(g / a_gdiv) * a_gs + m;
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += a[(int64_t)k * lda];
  out[(int64_t)g * out_gs + m] = s;
}

}  // namespace

bool tc_gemm_eligible(const GemmArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (a.M % BM != 0 || a.N != BN || a.K % BK != 0 || a.accumulate) return false;
  if (a.a_kmajor == 0 && a.b_kmajor != 0) return false;
  if (!al16(a.A) || !al16(a.B) || !al16(a.C) || a.lda % 4 || a.ldb % 4 || a.ldc % 4 || a.a_gs % 4 || a.b_gs % 4 || a.c_gs % 4) return false;
  if (a.bias && (!al16(a.bias) || a.bias_gs % 4)) return false;
  if (a.mask && (!al16(a.mask) || a.ldmask % 4 || a.mask_gs % 4)) return false;
  return true;
}

int tc_gemm_init() {
  IL_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return 0;
}

int launch_tc_gemm(il_handle* h, const GemmArgs& a, cudaStream_t stream) {
  IL_CHECK(tc_gemm_eligible(a), "tc_gemm: shape/layout not eligible (M=%d N=%d K=%d)", a.M, a.N, a.K);
  TcParams p;
  p.g = a;
  p.tiles_m = a.M / BM;
  p.n_tiles = a.G * p.tiles_m;
  p.split = h->gemm_mode == IL_GEMM_TF32X3 ? 1 : 0;
  const int grid = p.n_tiles < h->sm_count ? p.n_tiles : h->sm_count;
  IL_LAUNCH(h, tc_gemm_kernel, grid, THREADS, SMEM_BYTES, stream, p);
  if (a.colsum) {
    dim3 cg((a.M + 127) / 128, a.G);
    IL_LAUNCH(h, colsum_kernel, cg, 128, 0, stream, a.A, a.a_gs, a.a_gdiv, a.lda, a.K, a.M, a.colsum, a.colsum_gs);
  }
  return 0;
}
