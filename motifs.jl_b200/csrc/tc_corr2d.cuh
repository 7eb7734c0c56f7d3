// tcgen05 / TMEM path for the one genuinely dense contraction of the network at batched shapes:
//   U2 "corr2d":  out[n,i,k] = sum_{a<h} sum_{j<2M} A[n,i+a,j] F[a][j][k]      (model.jl:214,251)
// as the GEMM  [rows = NS*c] x [K = h*2M]  .  [K x 24]  with the im2col rows never materialised: row r's operand is the
// contiguous window A[r .. r+h-1][:] of the position-major buffer.
//
// Precision: operands are rounded to BF16, products accumulate in FP32 in TMEM (kind::f16).  This is the "stated bf16
// tolerance on tensor-core paths" mode of the north star: it is OFF by default (the fp32 SIMT kernels are the parity path) and is
// only wired into forward-only code retrieval (mb200_csc_create(..., forward_only = 2)).
//
// Layout trick that makes the sliding window free: the staged rows live in shared memory as [chunk of 8 columns][row][16 B].
// In the canonical K-major no-swizzle UMMA layout a core matrix is 8 rows x 16 B with the rows 16 B apart — exactly 8
// consecutive rows of one chunk — so the operand of window offset `a` is the SAME buffer with the descriptor start address
// advanced by a*16 B; SBO (next 8 rows) = 128 B, LBO (next K chunk) = R*16 B.  One staged copy of 128+h-1 rows serves all h
// offsets; nothing is re-read from L2 per offset.
#pragma once
#include <cuda_bf16.h>

#define TC_M 128            // output rows per tile = TMEM lanes
#define TC_N 32             // accumulator columns (K filters padded 24 -> 32)
#define TC_CH 14            // 16-byte chunks per staged row: 13 hold 2M=100 (+4 zero) columns, chunk 13 is all zero (pairs up chunk 12)

__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
    // [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type 0
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tc_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();             // never hang the device on a lost completion
    } while (!done);
}

// ---------------------------------------------------------------------------------------------------------------------
// Tap-grouped version (the one the tape launches when h % 4 == 0 and 4K is a legal UMMA N).
// With N = K = 24(32) accumulator columns every tcgen05.mma must fetch a 128 x 16 BF16 A operand (4 KB) from shared memory for
// only 16 cycles of math, so the tensor pipe waits on operand fetch (ncu, k_corr2d_tc2: sm__pipe_tensor_cycles_active 15.8 %).
// Here the h window offsets a = 4t + j are split into TC_J = 4 groups j that share one A operand:
//     D[r, (j,k)] = sum_{t < h/4} sum_ch A[r + 4t, ch] * F[4t + j][ch][k]            one MMA chain, N = 4K = 96, no padding columns
//     out[r, k]   = sum_{j < 4} D[r + j, (j,k)]                                      in the epilogue
// i.e. 4x fewer MMAs, each with 4x the math per fetched A byte (21 MMAs of 48 cycles per tile instead of 84 of 16).  The
// epilogue's row shift j is a warp shuffle (TMEM lane = row = thread); the three rows a warp needs from the next TMEM lane
// quarter cross through a small shared-memory buffer, and the last three rows of a tile belong to the next tile: tiles advance
// by 125 rows.
// ---------------------------------------------------------------------------------------------------------------------
#define TC_J 4
#define TC_VALID (TC_M - (TC_J - 1))      // output rows per tile
#define TC3_STAGES 4                      // shared-memory stages of the A operand (30.5 KB each)
#define TC3_THREADS 320                   // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 and 6-9 two epilogue groups

// F fp32 [h][2M][K] -> bf16 [h/4][TC_CH][4K][8]: row n = j*K + k of the B operand of tap step t holds F[4t + j][.][k]
__global__ void __launch_bounds__(256) k_tc_prep_F3(const float* __restrict__ F, __nv_bfloat16* __restrict__ Fb, int h, int M2, int K) { PDL_SYNC();
    const int N3 = TC_J * K;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (h / TC_J) * TC_CH * N3 * 8) return;
    const int e = t & 7, n = (t >> 3) % N3, ch = (t / (8 * N3)) % TC_CH, ts = t / (8 * N3 * TC_CH);
    const int j = n / K, k = n - j * K, col = ch * 8 + e;
    Fb[t] = __float2bfloat16_rn(col < M2 ? F[((int64_t)(TC_J * ts + j) * M2 + col) * K + k] : 0.f);
}

// A fp32 [rows][2M] -> bf16 chunk planes [TC_CH-1][arows][8]: plane ch holds columns 8ch..8ch+7 of every row, so the rows r0..r0+R-1 of
// one plane are one contiguous run of R*16 bytes — exactly one [chunk][row][16 B] slab of the shared-memory stage, fetched by a
// single 1-D bulk copy.  (A 3-D tensor-map box with a 16-byte inner extent, as k_corr2d_tc2 uses on the row-major buffer, turns
// every 16 bytes into a TMA request of its own: 1768 requests per tile, ~2 us, which was that kernel's real limit.)
__global__ void __launch_bounds__(256) k_tc_prep_A3(const float* __restrict__ A, __nv_bfloat16* __restrict__ Ab, int64_t rows, int64_t arows, int M2) { PDL_SYNC();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread = one (chunk, row) = 8 columns
    if (t >= rows * (TC_CH - 1)) return;
    const int ch = (int)(t / rows);
    const int64_t r = t - (int64_t)ch * rows;
    const float* src = A + r * M2 + ch * 8;
    __align__(16) __nv_bfloat16 v[8];
    #pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(ch * 8 + e < M2 ? src[e] : 0.f);
    *reinterpret_cast<uint4*>(Ab + ((int64_t)ch * arows + r) * 8) = *reinterpret_cast<const uint4*>(v);
}

template <int KK, int TT>
__global__ void __launch_bounds__(TC3_THREADS, 1) k_corr2d_tc3(const __nv_bfloat16* __restrict__ Ab, int64_t arows, const __nv_bfloat16* __restrict__ Fb,
                                                                float* __restrict__ out, int64_t rows_total, int ntiles, CscDims d) { PDL_SYNC();
    constexpr int N3 = TC_J * KK;                 // accumulator columns
    constexpr int ACC_STRIDE = 128;               // TMEM columns between the two accumulators
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    constexpr int T = TT;                         // window offsets per group = h / 4
    constexpr int R = TC_M + TC_J * (TT - 1);     // staged rows per tile
    constexpr uint32_t stage_bytes = (uint32_t)(((size_t)TC_CH * R * 16 + 1023) & ~(size_t)1023);
    uint8_t* sA0 = tc_smem;
    uint8_t* sB = tc_smem + TC3_STAGES * (size_t)stage_bytes;
    float* xch = reinterpret_cast<float*>(sB + (size_t)T * TC_CH * N3 * 16);      // [2 groups][2][3 warps][3 rows][3*KK]
    __shared__ __align__(8) uint64_t s_bars[2 * TC3_STAGES + 5];  // full[S], empty[S], accfull[2], accempty[2], B landed
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int i) { return (uint32_t)__cvta_generic_to_shared(&s_bars[i]); };

    if (tid == 0) {
        for (int i = 0; i < 2 * TC3_STAGES + 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar(i)) : "memory");
        for (int i = 2 * TC3_STAGES + 2; i < 2 * TC3_STAGES + 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" :: "r"(bar(i)) : "memory");   // 4 epilogue warps
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar(2 * TC3_STAGES + 4)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(2 * ACC_STRIDE) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // the zero chunk of every A stage (the resident B operand arrives by one bulk copy issued by the MMA thread below)
        for (int st = 0; st < TC3_STAGES; ++st) {
            uint4* z = reinterpret_cast<uint4*>(sA0 + (size_t)st * stage_bytes + (size_t)(TC_CH - 1) * R * 16);
            for (int i = tid; i < R; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t sA_addr = (uint32_t)__cvta_generic_to_shared(sA0), sB_addr = (uint32_t)__cvta_generic_to_shared(sB);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N3 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t tx_bytes = (uint32_t)(TC_CH - 1) * R * 16;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int st = it % TC3_STAGES; const uint32_t ph = (it / TC3_STAGES) & 1;
                tc_bar_wait(bar(TC3_STAGES + st), ph ^ 1);                               // stage free
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar(st)), "r"(tx_bytes) : "memory");
                const __nv_bfloat16* src = Ab + (int64_t)tile * TC_VALID * 8;
                for (int ch = 0; ch < TC_CH - 1; ++ch)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(sA_addr + st * stage_bytes + (uint32_t)ch * R * 16), "l"(src + (int64_t)ch * arows * 8), "r"((uint32_t)R * 16), "r"(bar(st)) : "memory");
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            {   // all of F (64.5 KB) in one bulk copy, overlapping the first A loads
                constexpr uint32_t b_bytes = (uint32_t)T * TC_CH * N3 * 16;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar(2 * TC3_STAGES + 4)), "r"(b_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(sB_addr), "l"(Fb), "r"(b_bytes), "r"(bar(2 * TC3_STAGES + 4)) : "memory");
                tc_bar_wait(bar(2 * TC3_STAGES + 4), 0);
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int st = it % TC3_STAGES; const uint32_t ph = (it / TC3_STAGES) & 1;
                const int ac = it & 1; const uint32_t aph = (it >> 1) & 1;
                tc_bar_wait(bar(st), ph);                                                // operands landed
                tc_bar_wait(bar(2 * TC3_STAGES + 2 + ac), aph ^ 1);                      // accumulator drained
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // the 21 descriptor pairs differ from the first by compile-time multiples of 16 bytes in the start-address field: the
                // loops unroll into one 32-bit add per descriptor, so the single issuing thread keeps ahead of the tensor pipe
                const uint64_t da0 = tc_desc(sA_addr + st * stage_bytes, (uint32_t)R * 16, 128);
                const uint64_t db0 = tc_desc(sB_addr, N3 * 16, 128);
                #pragma unroll
                for (int ts = 0; ts < T; ++ts) {
                    #pragma unroll
                    for (int t = 0; t < TC_CH / 2; ++t) {
                        const uint64_t da = da0 + (uint64_t)((2 * t) * R + TC_J * ts);
                        const uint64_t db = db0 + (uint64_t)((ts * TC_CH + 2 * t) * N3);
                        tc_mma_bf16(tmem + ac * ACC_STRIDE, da, db, idesc, (ts | t) ? 1u : 0u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar(TC3_STAGES + st)) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar(2 * TC3_STAGES + ac)) : "memory");
            }
        }
    } else {
        // two epilogue groups of four warps (one warp per TMEM lane quarter): group e drains accumulator e, i.e. every other tile,
        // so one tile's TMEM reads, shuffles and stores overlap the next tile's
        const int q = warp & 3;                                                           // TMEM lane quarter this warp may read
        const int eg = (warp - 2) >> 2;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;            // accumulator index and its phase
            if (st != eg) continue;
            tc_bar_wait(bar(2 * TC3_STAGES + st), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[N3];
            #pragma unroll
            for (int c0 = 0; c0 < N3; c0 += 8) {
                uint32_t u[8];
                const uint32_t taddr = tmem + st * ACC_STRIDE + c0 + ((uint32_t)(q * 32) << 16);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr) : "memory");
                #pragma unroll
                for (int i = 0; i < 8; ++i) v[c0 + i] = __uint_as_float(u[i]);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar(2 * TC3_STAGES + 2 + st)) : "memory");    // accumulator may be overwritten
            // rows 0..2 of lane quarters 1..3 go to the quarter below: [dst quarter][row][(j-1)*KK + k], 16-byte stores
            float4* xb = reinterpret_cast<float4*>(xch + (size_t)(eg * 2 + ((it >> 1) & 1)) * 3 * (TC_J - 1) * (TC_J - 1) * KK);
            constexpr int XROW = (TC_J - 1) * KK / 4;                                     // float4 per exchanged row
            if (q > 0 && lane < TC_J - 1) {
                float4* dst = xb + ((q - 1) * (TC_J - 1) + lane) * XROW;
                #pragma unroll
                for (int i = 0; i < XROW; ++i) dst[i] = make_float4(v[KK + 4 * i], v[KK + 4 * i + 1], v[KK + 4 * i + 2], v[KK + 4 * i + 3]);
            }
            asm volatile("bar.sync %0, 128;" :: "r"(1 + eg) : "memory");                // the four warps of this group
            float o[KK];
            #pragma unroll
            for (int k = 0; k < KK; ++k) o[k] = v[k];
            #pragma unroll
            for (int j = 1; j < TC_J; ++j) {
                // partner row = this row + j: lane + j of this warp, or row lane + j - 32 of the next quarter (one branch per j)
                const bool cross = lane + j >= 32;
                float xv[KK];
                #pragma unroll
                for (int k = 0; k < KK; ++k) xv[k] = 0.f;
                if (cross && q < 3) {
                    const float4* src = xb + (q * (TC_J - 1) + (lane + j - 32)) * XROW + (j - 1) * (KK / 4);
                    #pragma unroll
                    for (int i = 0; i < KK / 4; ++i) { const float4 w = src[i]; xv[4 * i] = w.x; xv[4 * i + 1] = w.y; xv[4 * i + 2] = w.z; xv[4 * i + 3] = w.w; }
                }
                #pragma unroll
                for (int k = 0; k < KK; ++k) {
                    const float sh = __shfl_down_sync(0xffffffffu, v[j * KK + k], j);
                    o[k] += cross ? xv[k] : sh;
                }
            }
            const int lr = q * 32 + lane;                                                // row of the tile
            const int64_t r = (int64_t)tile * TC_VALID + lr;
            if (lr < TC_VALID && r < rows_total) {
                const int64_t n = r / d.c;
                const int i = (int)(r - n * d.c);
                if (i < d.l) {
                    float4* op = reinterpret_cast<float4*>(out + (n * d.l + i) * KK);
                    #pragma unroll
                    for (int k = 0; k < KK / 4; ++k) op[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(2 * ACC_STRIDE) : "memory");
}
