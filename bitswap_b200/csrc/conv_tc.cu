// tcgen05 implicit-GEMM convolution for the dense W->W layers (sm_100a).
//
// Replaces F.conv2d inside WnConv2d._forward for the 3x3 / 5x5 W->W convolutions of the ResNet blocks
// (utils/torch/modules.py:98-106 called from ResNetLayer.forward :229-241), which are 98 % of the
// model's FLOPs.
//
// GEMM view (per image):  D[p, o] = sum_{tap, c} A_tap[p, c] * W[tap][o, c]
//     M = 256 pixels of one 16x16 image (two UMMA M=128 halves), N = 128 output channels per CTA
//     (grid = images x 2), K = taps x 256 input channels, walked as (tap, 32-channel chunk) k-blocks.
//   * A_tap is never materialised (no im2col): the activation tensor is NHWC bf16 and one 4-D TMA box
//     {32 ch, 16 w, 16 h, 1 img} at coordinates (c0, dx-r, dy-r, img) IS the shifted tile; TMA's
//     out-of-bounds zero fill supplies the "same" padding.
//   * float32 accuracy on bf16 tensor cores: every operand is split x = hi + lo (two bf16 planes) and
//     each k-block issues three MMAs: hi*hi into a MAIN float32 TMEM accumulator, hi*lo and lo*hi into a
//     separate CROSS accumulator.  The tensor core truncates (RZ) on every accumulate; keeping the 2^-8
//     smaller cross terms out of the main accumulator cuts its truncating adds 3x (measured: one shared
//     accumulator gave 1.9e-4 max error through the deepest stack, above the 1e-4 bar).
//   * accumulators = {main, cross} x 2 halves x (128 lanes x 128 columns) = all 512 TMEM columns.
//   * warp roles: warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane) and
//     TMEM allocator, warps 2..9 = epilogue (tcgen05.ld -> bias / residual / ELU / bf16 split -> HBM).
//   * fixed k order, no split-K, no atomics: results are bit-identical for any batch size or position
//     in the batch (the decoder must regenerate the encoder's tables exactly).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <vector>
#include "bsw_common.cuh"
#include "nets.cuh"

extern "C" int bsw_has_tensor_cores(void) { return 1; }

namespace {

constexpr int BK = 32;                      // channels per k-block (64-byte rows, SWIZZLE_64B)
constexpr int NSTAGE = 4;
constexpr int BN = 128;                     // output channels per CTA
constexpr int TILE_BYTES = 256 * BK * 2;    // one 256-row x 32-col bf16 A tile = 16 KB
constexpr int WTILE_BYTES = BN * BK * 2;    // one 128-row weight tile = 8 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * WTILE_BYTES; // A_hi, A_lo, B_hi, B_lo = 48 KB
constexpr int TC_THREADS = 320;             // 10 warps
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
// half-image variant (k_conv_tc_h): M = 128 pixels per CTA, 256 TMEM columns, 3 x 32 KB stages -> two CTAs per SM
constexpr int H_TILE_BYTES = 128 * BK * 2;  // 8 KB
constexpr int H_STAGE_BYTES = 2 * H_TILE_BYTES + 2 * WTILE_BYTES;   // 32 KB
constexpr int H_NSTAGE = 3;
constexpr int H_SMEM_BYTES = H_NSTAGE * H_STAGE_BYTES + 1024 + 256;

struct TcState {
    // bf16 hi/lo activation planes, ping-pong: [max_batch, 256 px, 256 ch]
    __nv_bfloat16 *act[2][2];
    CUtensorMap act_map[2][2];
    CUtensorMap act_map_h[2][2];   // same planes, box = 8 image rows (k_conv_tc_h)
    // in-conv input: the flat `given` re-laid as NHWC bf16 hi/lo planes with the (1..12) input channels padded to 32
    __nv_bfloat16 *inp[2];
    CUtensorMap inp_map[2];
    CUtensorMap inp_map_h[2];
    std::vector<void *> wbufs;
};

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed TMA/MMA must not hang the GPU -- trap instead (surfaces as a CUDA error).
// The bound is wall time (2 s on %globaltimer, sampled every 2^14 failed polls), not a spin count: a time-sliced or
// preempted context makes honest waits arbitrarily long in spins but not in device time spent inside this kernel.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3fffu) == 0) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) | [32,46) SBO >> 4 = 512 B
//   (8 rows x 64 B) | [46,48) version = 1 | [61,64) layout type 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ELU(x) = x > 0 ? x : exp(x) - 1 (utils/torch/modules.py:229-241 F.elu).  exp by the hardware ex2 (2 ulp of the result,
// |error| <= 2.4e-7 absolute after the subtraction): far inside the 1e-4 bar on mu/sigma (measured in the P3 tests) and
// a third of expm1f's instructions -- the epilogue's ALU work, not its bytes, is what the convs wait for.
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : __expf(x) - 1.0f; }

struct TcArgs {
    int taps, ks;
    int cchunks;             // 32-channel chunks per tap: 8 for the dense W->W convs, 1 for the in-convs
    const float *bias;       // [256]
    const float *resid;      // [n,256,256] fp32 or null
    float *T;                // trunk out or null
    int T_elu;
    __nv_bfloat16 *A_hi, *A_lo;   // next conv's input planes or null
    int A_elu;
};

__device__ __forceinline__ void conv_tc_body(const CUtensorMap &amap_hi, const CUtensorMap &amap_lo, const CUtensorMap &wmap_hi,
                                             const CUtensorMap &wmap_lo, const TcArgs &a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = (uint64_t *)(smem + NSTAGE * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + NSTAGE;
    uint64_t *acc_bar = empty_bar + NSTAGE;
    uint32_t *tmem_ptr = (uint32_t *)(acc_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x, co0 = blockIdx.y * BN;
    const int nkb = a.taps * a.cchunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_lo) : "memory");
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {     // TMEM: all 512 columns (2 accumulators of 256 fp32 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const int r = a.ks / 2;
            int stage = 0, phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                int tap = kb / a.cchunks, c0 = (kb - tap * a.cchunks) * BK;
                int dy = tap / a.ks, dx = tap - dy * a.ks;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t *st = smem + stage * STAGE_BYTES;
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                tma_load_4d(st, &amap_hi, &full_bar[stage], c0, dx - r, dy - r, img);
                tma_load_4d(st + TILE_BYTES, &amap_lo, &full_bar[stage], c0, dx - r, dy - r, img);
                tma_load_3d(st + 2 * TILE_BYTES, &wmap_hi, &full_bar[stage], c0, co0, tap);
                tma_load_3d(st + 2 * TILE_BYTES + WTILE_BYTES, &wmap_lo, &full_bar[stage], c0, co0, tap);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t aoff = half * (128 * BK * 2) + kk * 32;           // 128 rows x 64 B ; +32 B per UMMA_K
                        uint64_t a_hi = make_desc_sw64(sbase + aoff);
                        uint64_t a_lo = make_desc_sw64(sbase + TILE_BYTES + aoff);
                        uint64_t b_hi = make_desc_sw64(sbase + 2 * TILE_BYTES + kk * 32);
                        uint64_t b_lo = make_desc_sw64(sbase + 2 * TILE_BYTES + WTILE_BYTES + kk * 32);
                        uint32_t d_main = tmem_base + half * BN, d_cross = tmem_base + 256 + half * BN;
                        umma_bf16(d_main, a_hi, b_hi, (kb | kk) != 0);
                        umma_bf16(d_cross, a_lo, b_hi, (kb | kk) != 0);
                        umma_bf16(d_cross, a_hi, b_lo, 1);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_bar);                        // accumulators complete
        }
    } else {
        // ===== epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).., columns of accumulator (w-2)/4 =====
        // Coalesced through shared memory (see k_conv_tc_h): once acc_bar has fired every pipeline stage is idle, so each
        // warp transposes its 32 px x 32 ch chunks through a private padded tile in the stage memory and moves whole
        // 128-byte lines (tcgen05.ld hands a lane one pixel's 32 channels: written directly that is 32 scattered 16-byte
        // pieces per store instruction).  Same arithmetic, same order -> same bits as the direct epilogue.
        const int ew = warp - 2;
        const int half = ew >> 2;
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        float *tile = reinterpret_cast<float *>(smem) + ew * (32 * 36);          // 32 rows x 36 floats (144 B stride)
        const int rsub = lane >> 3, col4 = (lane & 7) * 4;
        const int64_t pbase = ((int64_t)img * 256 + half * 128 + quad * 32 + rsub) * 256 + co0 + col4;   // + (4 k) * 256 + cc
        // The residual is the trunk itself (updated in place: resid == T), so the compiler may not move its loads above the
        // stores of an earlier row -- left alone, every row's HBM latency is paid in series (ncu: the conv2 launches ran
        // 0.98 ms against 0.75 ms for conv1, long_scoreboard on the residual loads).  The loads are therefore issued
        // explicitly one 32-channel chunk ahead: chunk 0's before the accumulator is even ready (they do not depend on
        // it), chunk c+1's before chunk c's arithmetic and stores.  A thread re-reads nothing it has written (chunks are
        // disjoint channel ranges of its own rows), so the in-place update stays exact.
        float4 q[8], qn[8];
        auto load_resid = [&](float4 (&dst)[8], int cc) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                dst[k] = a.resid ? *reinterpret_cast<const float4 *>(a.resid + pbase + (int64_t)(4 * k) * 256 + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        load_resid(q, 0);
        mbar_wait(acc_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 32) {
            const int c0 = co0 + cc;
            {
                uint32_t rr[32], rc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + half * BN + cc, rr);
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + 256 + half * BN + cc, rc);
                float4 *trow = reinterpret_cast<float4 *>(tile + lane * 36);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    trow[i] = make_float4(__uint_as_float(rr[4 * i]) + __uint_as_float(rc[4 * i]),
                                          __uint_as_float(rr[4 * i + 1]) + __uint_as_float(rc[4 * i + 1]),
                                          __uint_as_float(rr[4 * i + 2]) + __uint_as_float(rc[4 * i + 2]),
                                          __uint_as_float(rr[4 * i + 3]) + __uint_as_float(rc[4 * i + 3]));
            }
            __syncwarp();
            if (cc + 32 < BN) load_resid(qn, cc + 32);
            const float4 bq = __ldg(reinterpret_cast<const float4 *>(a.bias + c0 + col4));
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = 4 * k + rsub;
                const int64_t prow = pbase + (int64_t)(4 * k) * 256 + cc;
                float4 x = *reinterpret_cast<const float4 *>(tile + r * 36 + col4);
                x.x += bq.x; x.y += bq.y; x.z += bq.z; x.w += bq.w;      // (main+cross)+bias
                if (a.resid) { x.x += q[k].x; x.y += q[k].y; x.z += q[k].z; x.w += q[k].w; }
                if (a.T_elu) { x.x = elu1(x.x); x.y = elu1(x.y); x.z = elu1(x.z); x.w = elu1(x.w); }
                if (a.T) *reinterpret_cast<float4 *>(a.T + prow) = x;
                if (a.A_hi) {
                    float y[4] = {x.x, x.y, x.z, x.w};
                    uint32_t hi[2], lo[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float x0 = a.A_elu ? elu1(y[2 * i]) : y[2 * i], x1 = a.A_elu ? elu1(y[2 * i + 1]) : y[2 * i + 1];
                        __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                        __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                        hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                    *reinterpret_cast<uint2 *>(a.A_hi + prow) = make_uint2(hi[0], hi[1]);
                    *reinterpret_cast<uint2 *>(a.A_lo + prow) = make_uint2(lo[0], lo[1]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = qn[k];
            __syncwarp();                                 // the tile is rewritten by the next chunk
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// One CTA per SM (shared memory); ~150 registers per thread for the epilogue's two residual chunks in flight.  That leaves
// 17 K registers per SM: room for the serial coder CTAs of another lane (11 K), not for a 32 K-register table CTA -- by
// measurement nothing is lost: the convs run at the board's power cap, so FP64 work beside them only lowers the clock
// (DESIGN.md 5.3).
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo,
          const __grid_constant__ CUtensorMap wmap_hi, const __grid_constant__ CUtensorMap wmap_lo, TcArgs a) {
    conv_tc_body(amap_hi, amap_lo, wmap_hi, wmap_lo, a);
}
// Half-image variant: the default for the in-convs (0.555 -> 0.416 ms per launch at B=1024), optional for the dense convs.
// k_conv_tc above owns the SM (192 KB of stages, all 512 TMEM columns), so the tensor
// pipe idles while its eight epilogue warps drain the accumulators to HBM -- about half of a 3x3 launch and nearly all
// of an in-conv launch (13.8 waves x ~40 us per CTA with nine k-blocks of work).  Here a CTA owns 128 pixels (8 image
// rows) x 128 output channels: {main, cross} x 128 columns = 256 TMEM columns and 3 x 32 KB stages, so TWO CTAs are
// resident per SM and one's epilogue runs under the other's MMAs.  Price: the weight tile is fetched per half image
// (L2->SM bytes per k-block and MMA row 48 KB/256 rows -> 32 KB/128 rows); the order of the k-blocks and of the MMAs
// into each accumulator is unchanged, so results are bit-identical to the full-tile kernel.
__global__ void __launch_bounds__(TC_THREADS, 2)
k_conv_tc_h(const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo,
            const __grid_constant__ CUtensorMap wmap_hi, const __grid_constant__ CUtensorMap wmap_lo, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = (uint64_t *)(smem + H_NSTAGE * H_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + H_NSTAGE;
    uint64_t *acc_bar = empty_bar + H_NSTAGE;
    uint32_t *tmem_ptr = (uint32_t *)(acc_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x, co0 = blockIdx.y * BN, mh = blockIdx.z;
    const int nkb = a.taps * a.cchunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_lo) : "memory");
        for (int s = 0; s < H_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {     // TMEM: 256 columns = main [0,128) + cross [128,256); the co-resident CTA takes the other 256
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {     // ===== TMA producer =====
            const int r = a.ks / 2;
            int stage = 0, phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                int tap = kb / a.cchunks, c0 = (kb - tap * a.cchunks) * BK;
                int dy = tap / a.ks, dx = tap - dy * a.ks;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t *st = smem + stage * H_STAGE_BYTES;
                mbar_expect_tx(&full_bar[stage], H_STAGE_BYTES);
                tma_load_4d(st, &amap_hi, &full_bar[stage], c0, dx - r, dy - r + 8 * mh, img);
                tma_load_4d(st + H_TILE_BYTES, &amap_lo, &full_bar[stage], c0, dx - r, dy - r + 8 * mh, img);
                tma_load_3d(st + 2 * H_TILE_BYTES, &wmap_hi, &full_bar[stage], c0, co0, tap);
                tma_load_3d(st + 2 * H_TILE_BYTES + WTILE_BYTES, &wmap_lo, &full_bar[stage], c0, co0, tap);
                if (++stage == H_NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {     // ===== MMA issuer =====
            int stage = 0, phase = 0;
            const uint32_t d_main = tmem_base, d_cross = tmem_base + BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                uint32_t sbase = smem_u32(smem + stage * H_STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    uint64_t a_hi = make_desc_sw64(sbase + kk * 32);
                    uint64_t a_lo = make_desc_sw64(sbase + H_TILE_BYTES + kk * 32);
                    uint64_t b_hi = make_desc_sw64(sbase + 2 * H_TILE_BYTES + kk * 32);
                    uint64_t b_lo = make_desc_sw64(sbase + 2 * H_TILE_BYTES + WTILE_BYTES + kk * 32);
                    umma_bf16(d_main, a_hi, b_hi, (kb | kk) != 0);
                    umma_bf16(d_cross, a_lo, b_hi, (kb | kk) != 0);
                    umma_bf16(d_cross, a_hi, b_lo, 1);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == H_NSTAGE) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_bar);
        }
    } else {
        // ===== epilogue: warp w may read TMEM lanes 32*(w%4)..; the two warps of a quadrant split the 128 columns =====
        const int ew = warp - 2;
        const int chalf = ew >> 2;
        const int quad = warp & 3;
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        // Coalesced epilogue.  tcgen05.ld hands every lane one pixel's 32 consecutive channels; written straight to
        // HBM that is 32 scattered 16-byte pieces per store instruction (the LSU, not HBM, bounds the direct epilogue:
        // ncu shows lg_throttle on every STG).  Instead the warp transposes each 32 px x 32 ch chunk through a private
        // padded tile in the (now idle) pipeline stages, so that 8 consecutive lanes hold 128 contiguous bytes of one
        // pixel row: residual loads, trunk stores and plane stores become 4 full lines per instruction.
        float *tile = reinterpret_cast<float *>(smem) + ew * (32 * 36);          // 32 rows x 36 floats (144 B stride)
        const int rsub = lane >> 3, col4 = (lane & 7) * 4;
#pragma unroll 1
        for (int cc = chalf * 64; cc < chalf * 64 + 64; cc += 32) {
            const int c0 = co0 + cc;
            {
                uint32_t rr[32], rc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + cc, rr);
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + BN + cc, rc);
                float4 *trow = reinterpret_cast<float4 *>(tile + lane * 36);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    trow[i] = make_float4(__uint_as_float(rr[4 * i]) + __uint_as_float(rc[4 * i]),
                                          __uint_as_float(rr[4 * i + 1]) + __uint_as_float(rc[4 * i + 1]),
                                          __uint_as_float(rr[4 * i + 2]) + __uint_as_float(rc[4 * i + 2]),
                                          __uint_as_float(rr[4 * i + 3]) + __uint_as_float(rc[4 * i + 3]));
            }
            __syncwarp();
            const float4 bq = __ldg(reinterpret_cast<const float4 *>(a.bias + c0 + col4));
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = 4 * k + rsub;
                const int64_t prow = ((int64_t)img * 256 + mh * 128 + quad * 32 + r) * 256 + c0 + col4;
                float4 x = *reinterpret_cast<const float4 *>(tile + r * 36 + col4);
                x.x += bq.x; x.y += bq.y; x.z += bq.z; x.w += bq.w;      // same order as the direct path: (main+cross)+bias
                if (a.resid) {
                    const float4 q = *reinterpret_cast<const float4 *>(a.resid + prow);
                    x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
                }
                if (a.T_elu) { x.x = elu1(x.x); x.y = elu1(x.y); x.z = elu1(x.z); x.w = elu1(x.w); }
                if (a.T) *reinterpret_cast<float4 *>(a.T + prow) = x;
                if (a.A_hi) {
                    float y[4] = {x.x, x.y, x.z, x.w};
                    uint32_t hi[2], lo[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float x0 = a.A_elu ? elu1(y[2 * i]) : y[2 * i], x1 = a.A_elu ? elu1(y[2 * i + 1]) : y[2 * i + 1];
                        __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                        __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                        hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                    *reinterpret_cast<uint2 *>(a.A_hi + prow) = make_uint2(hi[0], hi[1]);
                    *reinterpret_cast<uint2 *>(a.A_lo + prow) = make_uint2(lo[0], lo[1]);
                }
            }
            __syncwarp();                             // the tile is rewritten by the next chunk
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

// Persistent variant: ONE CTA per SM walks half-image tiles (128 pixels x 128 output channels) of the whole batch.
//   * TMEM ping-pong: tile i accumulates into columns [256*(i&1), +256) (main 128 + cross 128) while the eight epilogue
//     warps drain tile i-1 from the other 256 -- the epilogue (a third of k_conv_tc's CTA life on the 3x3 convs, nearly
//     all of it on the in-convs) disappears under the next tile's MMAs, and the TMA ring keeps prefetching across tile
//     boundaries, so the per-CTA start-up latency (allocation, tensormap fetch, first TMA round trip) is paid once per SM
//     instead of once per tile.
//   * Because the grid is at most one CTA per SM and every CTA is resident from the start, the block scheduler is free
//     to place other streams' kernels (the FP64-bound coder kernels of another lane: 32 K registers, no shared memory)
//     on the same SMs while the convolution runs -- with per-tile grids the queue of conv CTAs keeps them out.
//   * k-block order, MMA order per accumulator and the epilogue arithmetic are those of k_conv_tc / k_conv_tc_h: results
//     are bit-identical to both.
constexpr int P_NSTAGE = 5;
constexpr int P_SCRATCH = 8 * 32 * 36 * 4;   // one 32 x 36 float transposition tile per epilogue warp
constexpr int P_SMEM_BYTES = P_NSTAGE * H_STAGE_BYTES + P_SCRATCH + 1024 + 256;

// (min-blocks 2 only caps the registers at 96 per thread so that a 32 K-register coder CTA fits beside this one)
__global__ void __launch_bounds__(TC_THREADS, 2)
k_conv_tc_p(const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo,
            const __grid_constant__ CUtensorMap wmap_hi, const __grid_constant__ CUtensorMap wmap_lo, TcArgs a, int ntiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float *scratch = reinterpret_cast<float *>(smem + P_NSTAGE * H_STAGE_BYTES);
    uint64_t *full_bar = (uint64_t *)(smem + P_NSTAGE * H_STAGE_BYTES + P_SCRATCH);
    uint64_t *empty_bar = full_bar + P_NSTAGE;
    uint64_t *accf_bar = empty_bar + P_NSTAGE;           // [2] accumulator buffer complete (MMA -> epilogue)
    uint64_t *acce_bar = accf_bar + 2;                   // [2] accumulator buffer drained  (epilogue -> MMA)
    uint32_t *tmem_ptr = (uint32_t *)(acce_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = a.taps * a.cchunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_lo) : "memory");
        for (int s = 0; s < P_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&accf_bar[b], 1); mbar_init(&acce_bar[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {     // all 512 TMEM columns: two {main, cross} accumulator pairs
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // tile t -> image t>>2, output-channel half (t>>1)&1, pixel half t&1: the four tiles of an image run on neighbouring
    // SMs at about the same time, so the activation planes they share come out of L2.
    if (warp == 0) {
        if (lane == 0) {     // ===== TMA producer =====
            const int r = a.ks / 2;
            int stage = 0, phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int img = t >> 2, co0 = ((t >> 1) & 1) * BN, mh = t & 1;
                for (int kb = 0; kb < nkb; ++kb) {
                    int tap = kb / a.cchunks, c0 = (kb - tap * a.cchunks) * BK;
                    int dy = tap / a.ks, dx = tap - dy * a.ks;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + stage * H_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], H_STAGE_BYTES);
                    tma_load_4d(st, &amap_hi, &full_bar[stage], c0, dx - r, dy - r + 8 * mh, img);
                    tma_load_4d(st + H_TILE_BYTES, &amap_lo, &full_bar[stage], c0, dx - r, dy - r + 8 * mh, img);
                    tma_load_3d(st + 2 * H_TILE_BYTES, &wmap_hi, &full_bar[stage], c0, co0, tap);
                    tma_load_3d(st + 2 * H_TILE_BYTES + WTILE_BYTES, &wmap_lo, &full_bar[stage], c0, co0, tap);
                    if (++stage == P_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {     // ===== MMA issuer =====
            int stage = 0, phase = 0, it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&acce_bar[buf], ((it >> 1) & 1) ^ 1);      // the epilogue has drained this buffer (first use: passes)
                tc_fence_after();
                const uint32_t d_main = tmem_base + buf * 256, d_cross = d_main + BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint32_t sbase = smem_u32(smem + stage * H_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        uint64_t a_hi = make_desc_sw64(sbase + kk * 32);
                        uint64_t a_lo = make_desc_sw64(sbase + H_TILE_BYTES + kk * 32);
                        uint64_t b_hi = make_desc_sw64(sbase + 2 * H_TILE_BYTES + kk * 32);
                        uint64_t b_lo = make_desc_sw64(sbase + 2 * H_TILE_BYTES + WTILE_BYTES + kk * 32);
                        umma_bf16(d_main, a_hi, b_hi, (kb | kk) != 0);
                        umma_bf16(d_cross, a_lo, b_hi, (kb | kk) != 0);
                        umma_bf16(d_cross, a_hi, b_lo, 1);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == P_NSTAGE) { stage = 0; phase ^= 1; }
                }
                umma_commit(&accf_bar[buf]);
            }
        }
    } else {
        // ===== epilogue (the coalesced one of k_conv_tc_h): warp w may read TMEM lanes 32*(w%4)..; the two warps of a
        // quadrant split the 128 columns =====
        const int ew = warp - 2;
        const int chalf = ew >> 2;
        const int quad = warp & 3;
        float *tile = scratch + ew * (32 * 36);
        const int rsub = lane >> 3, col4 = (lane & 7) * 4;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int img = t >> 2, co0 = ((t >> 1) & 1) * BN, mh = t & 1;
            const int buf = it & 1;
            mbar_wait(&accf_bar[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t tb = tmem_base + buf * 256 + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int cc = chalf * 64; cc < chalf * 64 + 64; cc += 32) {
                const int c0 = co0 + cc;
                {
                    uint32_t rr[32], rc[32];
                    tmem_ld32(tb + cc, rr);
                    tmem_ld32(tb + BN + cc, rc);
                    if (cc == chalf * 64 + 32) {          // my last TMEM read of this tile: hand the buffer back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acce_bar[buf])) : "memory");
                    }
                    float4 *trow = reinterpret_cast<float4 *>(tile + lane * 36);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        trow[i] = make_float4(__uint_as_float(rr[4 * i]) + __uint_as_float(rc[4 * i]),
                                              __uint_as_float(rr[4 * i + 1]) + __uint_as_float(rc[4 * i + 1]),
                                              __uint_as_float(rr[4 * i + 2]) + __uint_as_float(rc[4 * i + 2]),
                                              __uint_as_float(rr[4 * i + 3]) + __uint_as_float(rc[4 * i + 3]));
                }
                __syncwarp();
                const float4 bq = __ldg(reinterpret_cast<const float4 *>(a.bias + c0 + col4));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = 4 * k + rsub;
                    const int64_t prow = ((int64_t)img * 256 + mh * 128 + quad * 32 + r) * 256 + c0 + col4;
                    float4 x = *reinterpret_cast<const float4 *>(tile + r * 36 + col4);
                    x.x += bq.x; x.y += bq.y; x.z += bq.z; x.w += bq.w;      // same order as the direct path: (main+cross)+bias
                    if (a.resid) {
                        const float4 q = *reinterpret_cast<const float4 *>(a.resid + prow);
                        x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
                    }
                    if (a.T_elu) { x.x = elu1(x.x); x.y = elu1(x.y); x.z = elu1(x.z); x.w = elu1(x.w); }
                    if (a.T) *reinterpret_cast<float4 *>(a.T + prow) = x;
                    if (a.A_hi) {
                        float y[4] = {x.x, x.y, x.z, x.w};
                        uint32_t hi[2], lo[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float x0 = a.A_elu ? elu1(y[2 * i]) : y[2 * i], x1 = a.A_elu ? elu1(y[2 * i + 1]) : y[2 * i + 1];
                            __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                            __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                            hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        *reinterpret_cast<uint2 *>(a.A_hi + prow) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint2 *>(a.A_lo + prow) = make_uint2(lo[0], lo[1]);
                    }
                }
                __syncwarp();                             // the tile is rewritten by the next chunk
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// Pair-tile variant: a 2-CTA cluster (one SM pair) owns one IMAGE and issues cta_group::2 MMAs with M = 256, N = 256.
//   * Why: the convs are bound by L2 -> SM delivery, ~6300 B/clk chip-wide = 42.6 B/clk/SM (B300_MICROARCH.md, "LTS
//     throughput cap", the same for TMA and LDG).  k_conv_tc fills 48 KB per k-block of twelve 64-cycle MMAs: 48 KB / 42.6 =
//     1154 cycles against a tensor floor of 768 -- measured 1170.  The half-image tiles fill 32 KB per six MMAs: 769 against
//     384 -- measured 768.  With N = 256 the activation tile is fetched once per 256 output channels, and with cta_group::2
//     the weight tile is SPLIT across the pair (128 of its 256 rows each): 32 KB per CTA and k-block of six 128-cycle MMAs
//     = 769 cycles of L2 against 768 of tensor pipe -- balanced; measured 855 (38 B/clk/SM, 0.90 of either bound).
//   * CTA r holds image rows 8r..8r+7 (its 128 pixels) of the A tile and weight rows 128r..128r+127 of the B tile; its
//     accumulators are 128 lanes x {main, cross} x 256 columns = all 512 TMEM columns: no ping-pong, so the epilogue
//     first empties TMEM into shared memory and registers, hands it back, and does its global traffic under the next tile.
//   * Leader (cluster rank 0) issues the MMAs and owns the `full` barriers (both CTAs' TMA loads signal them:
//     cp.async.bulk.tensor ... cta_group::2 with the barrier address' peer bit cleared); tcgen05.commit multicasts to both CTAs'
//     `empty` and `accf` barriers; the epilogue warps of both CTAs arrive on the leader's `acce` barrier.
//   * k order, MMA order per accumulator and the epilogue arithmetic are those of the other kernels: bit-identical results.
constexpr int P2_WTILE = 128 * BK * 2;                         // my 128 of the tile's 256 weight rows: 8 KB per plane
constexpr int P2_STAGE = 2 * H_TILE_BYTES + 2 * P2_WTILE;      // 32 KB
constexpr int P2_NSTAGE = 4;
constexpr int P2_NE = 8;                                       // epilogue warps per CTA: 2 per TMEM lane quadrant, 128 columns each
constexpr int P2_SCRATCH = P2_NE * 3 * 32 * 32 * 4;            // three 32 x 32 float tiles per epilogue warp (96 KB): see the epilogue
constexpr int P2_SMEM_BYTES = P2_NSTAGE * P2_STAGE + P2_SCRATCH + 1024 + 256;
constexpr uint32_t IDESC_2SM = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((256u >> 4) << 24);   // M = 256, N = 256
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;                    // clears the bit that tells the pair's two shared windows apart

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of the pair: data into MY shared memory, completion bytes onto the LEADER's mbarrier
__device__ __forceinline__ void tma2_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(0x1000000000000000ull) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "l"(0x1000000000000000ull) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC_2SM), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t *bar) {       // arrives on BOTH CTAs' barrier at this offset
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) { // arrive on the leader CTA's barrier at this offset
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((2 + P2_NE) * 32, 1)
k_conv_tc_2sm(const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo,
              const __grid_constant__ CUtensorMap wmap_hi, const __grid_constant__ CUtensorMap wmap_lo, TcArgs a, int ntiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float *scratch = reinterpret_cast<float *>(smem + P2_NSTAGE * P2_STAGE);
    uint64_t *full_bar = (uint64_t *)(smem + P2_NSTAGE * P2_STAGE + P2_SCRATCH);
    uint64_t *empty_bar = full_bar + P2_NSTAGE;
    uint64_t *accf_bar = empty_bar + P2_NSTAGE;          // [2] accumulator buffer complete (leader's MMAs -> both epilogues)
    uint64_t *acce_bar = accf_bar + 2;                   // [2] accumulator buffer drained  (both epilogues -> leader's MMA thread)
    uint32_t *tmem_ptr = (uint32_t *)(acce_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_rank();                // 0 = leader
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nkb = a.taps * a.cchunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_lo) : "memory");
        for (int s = 0; s < P2_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accf_bar[0], 1); mbar_init(&acce_bar[0], 2 * P2_NE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {     // the pair's TMEM: 512 columns in both CTAs (same warp in both CTAs, same shared offset)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                  // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {     // ===== TMA producer (both CTAs): my half of A, my 64 weight rows =====
            const int r = a.ks / 2;
            int stage = 0, phase = 0;
            for (int t = pair; t < ntiles; t += npairs) {
                const int img = t;
                for (int kb = 0; kb < nkb; ++kb) {
                    int tap = kb / a.cchunks, c0 = (kb - tap * a.cchunks) * BK;
                    int dy = tap / a.ks, dx = tap - dy * a.ks;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + stage * P2_STAGE;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * P2_STAGE);       // both CTAs' bytes land on the leader's barrier
                    tma2_load_4d(st, &amap_hi, &full_bar[stage], c0, dx - r, dy - r + 8 * rank, img);
                    tma2_load_4d(st + H_TILE_BYTES, &amap_lo, &full_bar[stage], c0, dx - r, dy - r + 8 * rank, img);
                    tma2_load_3d(st + 2 * H_TILE_BYTES, &wmap_hi, &full_bar[stage], c0, 128 * rank, tap);
                    tma2_load_3d(st + 2 * H_TILE_BYTES + P2_WTILE, &wmap_lo, &full_bar[stage], c0, 128 * rank, tap);
                    if (++stage == P2_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {     // ===== MMA issuer: the leader, for the pair =====
            int stage = 0, phase = 0, it = 0;
            for (int t = pair; t < ntiles; t += npairs, ++it) {
                mbar_wait(&acce_bar[0], (it & 1) ^ 1);               // both CTAs' epilogues have drained the accumulators (first tile: passes)
                tc_fence_after();
                const uint32_t d_main = tmem_base, d_cross = tmem_base + 256;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint32_t sbase = smem_u32(smem + stage * P2_STAGE);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        uint64_t a_hi = make_desc_sw64(sbase + kk * 32);
                        uint64_t a_lo = make_desc_sw64(sbase + H_TILE_BYTES + kk * 32);
                        uint64_t b_hi = make_desc_sw64(sbase + 2 * H_TILE_BYTES + kk * 32);
                        uint64_t b_lo = make_desc_sw64(sbase + 2 * H_TILE_BYTES + P2_WTILE + kk * 32);
                        umma2_bf16(d_main, a_hi, b_hi, (kb | kk) != 0);
                        umma2_bf16(d_cross, a_lo, b_hi, (kb | kk) != 0);
                        umma2_bf16(d_cross, a_hi, b_lo, 1);
                    }
                    umma2_commit(&empty_bar[stage]);                 // frees the stage in both CTAs when these MMAs retire
                    if (++stage == P2_NSTAGE) { stage = 0; phase ^= 1; }
                }
                umma2_commit(&accf_bar[0]);
            }
        }
    } else {
        // ===== epilogue (both CTAs, my 128 pixels x 256 channels) =====
        // Phase 1 (exposed, short): main + cross summed out of TMEM -- three of my four 32-channel chunks into my shared
        // tiles, the fourth kept in registers -- then the accumulators go back to the leader's MMA thread.
        // Phase 2 (under the next tile's MMAs): bias, residual (loaded one chunk ahead), ELU, bf16 hi/lo split and the
        // coalesced global stores, chunk by chunk out of shared memory; the register chunk goes through tile 0 once that is done.
        // float4 i of row r sits at slot i ^ (r & 7) of its 128-byte row: conflict-free both ways without padding, which is
        // what lets 96 KB of tiles sit next to four 32 KB stages.  (With the whole 128 KB output tile in shared memory only
        // three stages fit, and the main loop slowed by more than the epilogue saved: profiles/r2_ab_runs.md.)
        constexpr int CW = 256 / (P2_NE / 4);            // columns per warp: 128 = four chunks
        static_assert(CW == 128, "epilogue written for four chunks per warp");
        const int ew = warp - 2;
        const int cpart = ew >> 2;                       // which CW of the 256 columns
        const int quad = warp & 3;
        float *slice = scratch + ew * (3 * 32 * 32);
        const int rsub = lane >> 3, col4 = (lane & 7) * 4;
        int it = 0;
        for (int t = pair; t < ntiles; t += npairs, ++it) {
            const int img = t;
            const int64_t pbase = ((int64_t)img * 256 + rank * 128 + quad * 32 + rsub) * 256 + cpart * CW + col4;   // + (4 k) * 256 + cc
            float4 q[8], qn[8];
            float s3[32];
            auto load_resid = [&](float4 (&dst)[8], int cc) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    dst[k] = a.resid ? *reinterpret_cast<const float4 *>(a.resid + pbase + (int64_t)(4 * k) * 256 + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            load_resid(q, 0);
            mbar_wait(&accf_bar[0], it & 1);
            tc_fence_after();
            const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + cpart * CW;
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                uint32_t rr[32], rc[32];
                tmem_ld32(tb + c * 32, rr);
                tmem_ld32(tb + 256 + c * 32, rc);
                float *trow = slice + c * 1024 + lane * 32;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4 *>(trow + ((i ^ (lane & 7)) << 2)) =
                        make_float4(__uint_as_float(rr[4 * i]) + __uint_as_float(rc[4 * i]),
                                    __uint_as_float(rr[4 * i + 1]) + __uint_as_float(rc[4 * i + 1]),
                                    __uint_as_float(rr[4 * i + 2]) + __uint_as_float(rc[4 * i + 2]),
                                    __uint_as_float(rr[4 * i + 3]) + __uint_as_float(rc[4 * i + 3]));
            }
            {
                uint32_t rr[32], rc[32];
                tmem_ld32(tb + 96, rr);
                tmem_ld32(tb + 256 + 96, rc);
#pragma unroll
                for (int i = 0; i < 32; ++i) s3[i] = __uint_as_float(rr[i]) + __uint_as_float(rc[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acce_bar[0]);     // the pair's next tile may start
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int cc = c * 32, c0 = cpart * CW + cc;
                const float *tl = slice + (c < 3 ? c : 0) * 1024;
                if (c == 3) {                             // the register chunk takes over tile 0 (its readers are done: same warp)
                    __syncwarp();
                    float *trow = slice + lane * 32;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4 *>(trow + ((i ^ (lane & 7)) << 2)) = make_float4(s3[4 * i], s3[4 * i + 1], s3[4 * i + 2], s3[4 * i + 3]);
                    __syncwarp();
                }
                if (c < 3) load_resid(qn, cc + 32);
                const float4 bq = __ldg(reinterpret_cast<const float4 *>(a.bias + c0 + col4));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = 4 * k + rsub;
                    const int64_t prow = pbase + (int64_t)(4 * k) * 256 + cc;
                    float4 x = *reinterpret_cast<const float4 *>(tl + r * 32 + (((lane & 7) ^ (r & 7)) << 2));
                    x.x += bq.x; x.y += bq.y; x.z += bq.z; x.w += bq.w;      // (main+cross)+bias
                    if (a.resid) { x.x += q[k].x; x.y += q[k].y; x.z += q[k].z; x.w += q[k].w; }
                    if (a.T_elu) { x.x = elu1(x.x); x.y = elu1(x.y); x.z = elu1(x.z); x.w = elu1(x.w); }
                    if (a.T) *reinterpret_cast<float4 *>(a.T + prow) = x;
                    if (a.A_hi) {
                        float y[4] = {x.x, x.y, x.z, x.w};
                        uint32_t hi[2], lo[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float x0 = a.A_elu ? elu1(y[2 * i]) : y[2 * i], x1 = a.A_elu ? elu1(y[2 * i + 1]) : y[2 * i + 1];
                            __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                            __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                            hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        *reinterpret_cast<uint2 *>(a.A_hi + prow) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint2 *>(a.A_lo + prow) = make_uint2(lo[0], lo[1]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) q[k] = qn[k];
            }
            __syncwarp();                                 // the tiles are rewritten by my next tile
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();                                  // the peer may signal my barriers / read my shared memory until here
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// mu/sigma heads on the tensor cores: the 3x3 W -> (n_mu + n_sc <= 32) convolution that ends every net
// (cifar_train.py:349,368,411,426), with the scale transforms and the UnSqueeze2d index map of the x head in the epilogue.
// GEMM per image: M = 256 pixels (two UMMA M=128 halves), N = 32 (output channels zero-padded), K = taps x 256.
// {main, cross} x 2 halves x 32 columns = 128 TMEM columns and 3 x 36 KB stages, so two CTAs share an SM and one's
// epilogue / start-up runs under the other's MMAs.  Narrow-N MMAs are bound by the A-operand bytes (shared memory reads
// and L2 -> SM), not by the tensor pipe: the point is to stop spending 16 ms per step in FP32 FMAs (k_conv_simt).
constexpr int HD_N = 32;
constexpr int HD_WTILE = HD_N * BK * 2;                       // 2 KB
constexpr int HD_STAGE = 2 * TILE_BYTES + 2 * HD_WTILE;        // 36 KB
constexpr int HD_NSTAGE = 3;
constexpr int HD_SMEM_BYTES = HD_NSTAGE * HD_STAGE + 1024 + 256;
constexpr uint32_t IDESC_HD = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HD_N >> 3) << 17) | ((128u >> 4) << 24);

struct HeadArgs {
    int taps, ks;
    const float *bias;       // [>= n_mu + n_sc]
    float *mu, *scale;
    int n_mu, n_sc, scale_kind, out_mode, out_dim;
};

__device__ __forceinline__ float hd_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float hd_softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }   // modules.py:112-114
__device__ __forceinline__ float hd_scale(float r, int kind) {      // same expressions as nets.cu scale_transform
    switch (kind) {
        case SCALE_INFER:   return 0.1f + 0.9f * hd_sigmoid(r + 2.f);                               // cifar_train.py:349,368
        case SCALE_DEEPGEN: return 0.1f + 0.9f * hd_softplus(r + 0.54132485461291810f);             // :426
        case SCALE_X:       return ((2.f / 255.f) / 8.f) + hd_softplus(r);                          // :411 / imagenetcrop :417
    }
    return r;
}

__global__ void __launch_bounds__(TC_THREADS, 2)
k_conv_tc_head(const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo,
               const __grid_constant__ CUtensorMap wmap_hi, const __grid_constant__ CUtensorMap wmap_lo, HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = (uint64_t *)(smem + HD_NSTAGE * HD_STAGE);
    uint64_t *empty_bar = full_bar + HD_NSTAGE;
    uint64_t *acc_bar = empty_bar + HD_NSTAGE;
    uint32_t *tmem_ptr = (uint32_t *)(acc_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x;
    const int nkb = a.taps * 8;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap_lo) : "memory");
        for (int s = 0; s < HD_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {     // 128 columns: main [0,64) = two halves x 32, cross [64,128)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {     // ===== TMA producer =====
            const int r = a.ks / 2;
            int stage = 0, phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                int tap = kb >> 3, c0 = (kb & 7) * BK;
                int dy = tap / a.ks, dx = tap - dy * a.ks;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t *st = smem + stage * HD_STAGE;
                mbar_expect_tx(&full_bar[stage], HD_STAGE);
                tma_load_4d(st, &amap_hi, &full_bar[stage], c0, dx - r, dy - r, img);
                tma_load_4d(st + TILE_BYTES, &amap_lo, &full_bar[stage], c0, dx - r, dy - r, img);
                tma_load_3d(st + 2 * TILE_BYTES, &wmap_hi, &full_bar[stage], c0, 0, tap);
                tma_load_3d(st + 2 * TILE_BYTES + HD_WTILE, &wmap_lo, &full_bar[stage], c0, 0, tap);
                if (++stage == HD_NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {     // ===== MMA issuer =====
            int stage = 0, phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                uint32_t sbase = smem_u32(smem + stage * HD_STAGE);
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t aoff = half * (128 * BK * 2) + kk * 32;
                        uint64_t a_hi = make_desc_sw64(sbase + aoff);
                        uint64_t a_lo = make_desc_sw64(sbase + TILE_BYTES + aoff);
                        uint64_t b_hi = make_desc_sw64(sbase + 2 * TILE_BYTES + kk * 32);
                        uint64_t b_lo = make_desc_sw64(sbase + 2 * TILE_BYTES + HD_WTILE + kk * 32);
                        uint32_t d_main = tmem_base + half * HD_N, d_cross = tmem_base + 64 + half * HD_N;
                        const uint32_t acc = (kb | kk) != 0;
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(d_main), "l"(a_hi), "l"(b_hi), "r"(IDESC_HD), "r"(acc) : "memory");
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(d_cross), "l"(a_lo), "l"(b_hi), "r"(IDESC_HD), "r"(acc) : "memory");
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(d_cross), "l"(a_hi), "l"(b_lo), "r"(IDESC_HD), "r"(1u) : "memory");
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == HD_NSTAGE) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_bar);
        }
    } else {
        // ===== epilogue: warp -> (pixel half, TMEM lane quadrant); a lane owns one pixel and all 32 output channels.
        // For a fixed channel the 32 lanes write 32 consecutive pixels of the flat CHW mu / sigma rows: whole lines. =====
        const int ew = warp - 2;
        const int half = ew >> 2;
        const int quad = warp & 3;
        const int p = half * 128 + quad * 32 + lane;
        const int y = p >> 4, x = p & 15;
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        uint32_t rr[32], rc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + half * HD_N, rr);
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + 64 + half * HD_N, rc);
        float *mu_row = a.mu + (int64_t)img * a.out_dim;
        float *sc_row = a.scale ? a.scale + (int64_t)img * a.out_dim : nullptr;
#pragma unroll
        for (int c = 0; c < HD_N; ++c) {
            if (c >= a.n_mu + a.n_sc) break;
            const bool is_mu = c < a.n_mu;
            const int o = is_mu ? c : c - a.n_mu;
            float r = (__uint_as_float(rr[c]) + __uint_as_float(rc[c])) + __ldg(a.bias + c);
            if (!is_mu) r = hd_scale(r, a.scale_kind);
            float *dst = is_mu ? mu_row : sc_row;
            if (a.out_mode == OUT_HEAD_Z) dst[o * 256 + p] = r;
            else dst[(o >> 2) * 1024 + (2 * y + ((o >> 1) & 1)) * 32 + 2 * x + (o & 1)] = r;      // UnSqueeze2d, modules.py:205-207
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    }
}

// float32 NHWC -> bf16 hi/lo planes (used once per net, after the SIMT in-conv)
__global__ void k_split_planes(const float *__restrict__ in, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = in[i];
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library must load (and export its
// symbols) on machines without libcuda.so.1, e.g. the CPU-only build container.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// flat CHW `given` (float32) -> NHWC bf16 hi/lo planes with 32 channels (zero padded).  mode IN_CHW_Z: channel c at
// given[c*256 + p]; IN_CHW_X: Squeeze2d(2) on the fly, channel c*4+fh*2+fw <- pixel (2h+fh, 2w+fw) of image channel c
// (utils/torch/modules.py:183-185).
__global__ void k_given_to_planes(const float *__restrict__ given, int in_dim, int cin, int mode, __nv_bfloat16 *__restrict__ hi,
                                  __nv_bfloat16 *__restrict__ lo, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (image, pixel, channel)
    if (i >= n * 256 * 32) return;
    int ch = (int)(i & 31);
    int p = (int)((i >> 5) & 255);
    int64_t img = i >> 13;
    float x = 0.f;
    if (ch < cin) {
        const float *src = given + img * in_dim;
        int y = p >> 4, xx = p & 15;
        x = (mode == IN_CHW_Z) ? src[ch * 256 + p] : src[(ch >> 2) * 1024 + (2 * y + ((ch >> 1) & 1)) * 32 + 2 * xx + (ch & 1)];
    }
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

int encode_map(CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
               const cuuint32_t *box) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        BSW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) {
            bsw_set_error("cuTensorMapEncodeTiled not available from this driver");
            return BSW_E_CUDA;
        }
        fn = (EncodeTiledFn)p;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        bsw_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return BSW_E_CUDA;
    }
    return BSW_OK;
}

}  // namespace

struct TcSlot { CUtensorMap map_hi, map_lo; };

int bsw_model_tc_prepare(bsw_model *m) {
    BSW_REQUIRE(m->Wp == 256, "tcgen05 conv path needs a padded width of 256 (reswidth 193..256)");
    TcState *ts = new TcState();
    m->tc_state = ts;
    const size_t act_elems = (size_t)m->d.max_batch * 256 * 256;
    for (int b = 0; b < 2; ++b)
        for (int pl = 0; pl < 2; ++pl) {
            BSW_CUDA(cudaMalloc(&ts->act[b][pl], act_elems * sizeof(__nv_bfloat16)));
            BSW_CUDA(cudaMemset(ts->act[b][pl], 0, act_elems * sizeof(__nv_bfloat16)));
            cuuint64_t dims[4] = {256, 16, 16, (cuuint64_t)m->d.max_batch};
            cuuint64_t str[3] = {256 * 2, 16 * 256 * 2, 256 * 256 * 2};
            cuuint32_t box[4] = {BK, 16, 16, 1};
            if (int rc = encode_map(&ts->act_map[b][pl], ts->act[b][pl], 4, dims, str, box)) return rc;
            cuuint32_t boxh[4] = {BK, 16, 8, 1};
            if (int rc = encode_map(&ts->act_map_h[b][pl], ts->act[b][pl], 4, dims, str, boxh)) return rc;
        }
    for (int pl = 0; pl < 2; ++pl) {
        size_t elems = (size_t)m->d.max_batch * 256 * 32;
        BSW_CUDA(cudaMalloc(&ts->inp[pl], elems * sizeof(__nv_bfloat16)));
        cuuint64_t dims[4] = {32, 16, 16, (cuuint64_t)m->d.max_batch};
        cuuint64_t str[3] = {32 * 2, 16 * 32 * 2, 256 * 32 * 2};
        cuuint32_t box[4] = {BK, 16, 16, 1};
        if (int rc = encode_map(&ts->inp_map[pl], ts->inp[pl], 4, dims, str, box)) return rc;
        cuuint32_t boxh[4] = {BK, 16, 8, 1};
        if (int rc = encode_map(&ts->inp_map_h[pl], ts->inp[pl], 4, dims, str, boxh)) return rc;
    }
    // weights: [tap][o=256][c=Kc] bf16 hi/lo planes, K (=c) innermost; Kc = 256 for the dense convs, 32 for the in-convs
    std::vector<TcSlot> *slots = new std::vector<TcSlot>();
    std::vector<char> is_in(m->convs.size(), 0), is_head(m->convs.size(), 0);
    for (auto &np : m->infer) { is_in[np.in_conv] = 1; is_head[np.head] = 1; }
    for (auto &np : m->gen) { is_in[np.in_conv] = 1; is_head[np.head] = 1; }
    for (size_t ci_ = 0; ci_ < m->convs.size(); ++ci_) {
        ConvSlot &c = m->convs[ci_];
        bool dense = (c.Cin == m->d.reswidth && c.Cout == m->d.reswidth);
        bool inconv = is_in[ci_] && c.Cin <= 32 && c.Cout == m->d.reswidth;
        bool head = is_head[ci_] && !dense && c.Cin == m->d.reswidth && c.Cout <= HD_N;
        if (!dense && !inconv && !head) continue;
        BSW_REQUIRE(!c.host_w.empty(), "tc_prepare: host weights already released");
        const int taps = c.ks * c.ks;
        const int Kc = inconv ? 32 : 256;
        const int rows = head ? HD_N : 256;              // output-channel rows per tap (heads: zero padded to 32)
        std::vector<__nv_bfloat16> hi((size_t)taps * rows * Kc), lo(hi.size());
        for (int tp = 0; tp < taps; ++tp)
            for (int ci = 0; ci < Kc; ++ci)
                for (int o = 0; o < rows; ++o) {
                    float w = (ci < c.CinP && o < c.CoutP) ? c.host_w[((size_t)tp * c.CinP + ci) * c.CoutP + o] : 0.f;
                    __nv_bfloat16 h = __float2bfloat16_rn(w);
                    size_t idx = ((size_t)tp * rows + o) * Kc + ci;
                    hi[idx] = h;
                    lo[idx] = __float2bfloat16_rn(w - __bfloat162float(h));
                }
        BSW_CUDA(cudaMalloc(&c.w_hi, hi.size() * 2));
        BSW_CUDA(cudaMalloc(&c.w_lo, lo.size() * 2));
        BSW_CUDA(cudaMemcpy(c.w_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
        BSW_CUDA(cudaMemcpy(c.w_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
        ts->wbufs.push_back(c.w_hi);
        ts->wbufs.push_back(c.w_lo);
        TcSlot s;
        cuuint64_t dims[3] = {(cuuint64_t)Kc, (cuuint64_t)rows, (cuuint64_t)taps};
        cuuint64_t str[2] = {(cuuint64_t)Kc * 2, (cuuint64_t)rows * Kc * 2};
        cuuint32_t box[3] = {BK, (cuuint32_t)(head ? HD_N : BN), 1};
        if (int rc = encode_map(&s.map_hi, c.w_hi, 3, dims, str, box)) return rc;
        if (int rc = encode_map(&s.map_lo, c.w_lo, 3, dims, str, box)) return rc;
        c.tc_index = (int)slots->size();
        slots->push_back(s);
    }
    m->tc_slots = slots;
    BSW_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BSW_CUDA(cudaFuncSetAttribute(k_conv_tc_h, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_BYTES));
    BSW_CUDA(cudaFuncSetAttribute(k_conv_tc_p, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    BSW_CUDA(cudaFuncSetAttribute(k_conv_tc_2sm, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES));
    bsw_prefer_max_shared(k_conv_tc_2sm);
    BSW_CUDA(cudaFuncSetAttribute(k_conv_tc_head, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_SMEM_BYTES));
    bsw_prefer_max_shared(k_conv_tc_head);
    // one SM-wide L1/shared split for every kernel of the path (see bsw_prefer_max_shared)
    bsw_prefer_max_shared(k_conv_tc); bsw_prefer_max_shared(k_conv_tc_h);
    bsw_prefer_max_shared(k_conv_tc_p); bsw_prefer_max_shared(k_given_to_planes);
    m->tc_ready = true;
    return BSW_OK;
}

void bsw_model_tc_release(bsw_model *m) {
    TcState *ts = (TcState *)m->tc_state;
    if (!ts) return;
    for (int b = 0; b < 2; ++b)
        for (int pl = 0; pl < 2; ++pl) cudaFree(ts->act[b][pl]);
    for (int pl = 0; pl < 2; ++pl) cudaFree(ts->inp[pl]);
    for (void *p : ts->wbufs) cudaFree(p);
    delete ts;
    delete (std::vector<TcSlot> *)m->tc_slots;
    m->tc_state = nullptr;
    m->tc_slots = nullptr;
}

// Plane buffer addressing for nets.cu: which == 0 -> the "A" ping-pong buffer, 1 -> "B".
void bsw_tc_planes(bsw_model *m, int which, void **hi, void **lo) {
    TcState *ts = (TcState *)m->tc_state;
    *hi = ts->act[which][0];
    *lo = ts->act[which][1];
}

int bsw_tc_split(bsw_model *m, const float *in, int which, int64_t n, cudaStream_t st) {
    TcState *ts = (TcState *)m->tc_state;
    int64_t cnt = n * 256 * 256;
    k_split_planes<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(in, ts->act[which][0], ts->act[which][1], cnt);
    BSW_LAUNCH_CHECK();
    return BSW_OK;
}

// Which kernel the dense convs and in-convs run on: a bit mask -- 8 / 16 = dense 3x3 / 5x5 on the pair-tile kernel
// k_conv_tc_2sm, 1 / 2 = dense 3x3 / 5x5 on the persistent half-image kernel k_conv_tc_p, 4 = in-convs on k_conv_tc_p; a
// clear bit means the per-tile grid (k_conv_tc / k_conv_tc_h).  Default 24 (BSW_TC_PERSIST overrides): measured on B200 at
// 1024 images the pair tiles win (3x3 0.83 -> 0.68 ms, 5x5 1.94 -> 1.55 ms) because the convs are bound by L2 -> SM bytes
// per MMA, which N = 256 and the weight-tile split halve; the half-image tiles of k_conv_tc_p lose to the per-tile grids
// for the same reason.  Results are bit-identical in every mode.
static int g_tc_mode = -1;
static bool tc_persistent(int which) {
    static const int env = getenv("BSW_TC_PERSIST") ? atoi(getenv("BSW_TC_PERSIST")) : 24;
    return (((g_tc_mode < 0) ? env : g_tc_mode) & which) != 0;
}
extern "C" int bsw_set_conv_mode(int mode) {
    BSW_REQUIRE(mode >= -1 && mode <= 31, "bsw_set_conv_mode: -1 (default) or a mask of 1 | 2 | 4 (persistent: dense 3x3, 5x5, in-convs) | 8 | 16 (pair tiles: dense 3x3, 5x5)");
    g_tc_mode = mode;
    return BSW_OK;
}
static unsigned tc_pgrid(int ntiles) {             // one CTA per SM of the current device, at most one per tile
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = v > 0 ? v : 148;
    }
    return (unsigned)(ntiles < sms[dev] ? ntiles : sms[dev]);
}

// a.in must be one of the plane buffers: a.in_planes = 0/1 selects it; outputs a.A_planes = 0/1/-1.
int bsw_conv_tc(bsw_model *m, const ConvSlot &c, const ConvArgs &a, int64_t n, cudaStream_t st) {
    TcState *ts = (TcState *)m->tc_state;
    BSW_REQUIRE(ts && c.tc_index >= 0, "bsw_conv_tc: conv has no tensor-core weights");
    const TcSlot &s = (*(std::vector<TcSlot> *)m->tc_slots)[c.tc_index];
    TcArgs t;
    t.taps = c.ks * c.ks; t.ks = c.ks;
    t.cchunks = 8;
    t.bias = c.bias; t.resid = a.resid; t.T = a.T; t.T_elu = a.T_elu;
    t.A_hi = a.A_planes >= 0 ? ts->act[a.A_planes][0] : nullptr;
    t.A_lo = a.A_planes >= 0 ? ts->act[a.A_planes][1] : nullptr;
    t.A_elu = a.A_elu;
    // Kernel choice by measurement (B200, 1024 images), see the mode mask above.  A 2-CTA cluster with multicast activation
    // tiles but M = 128 MMAs (r1) measured no gain (0.897 vs 0.880 ms) and is gone; so is a pair kernel with N = 128 and a
    // TMEM ping-pong (0.92 / 2.4 ms).
    if (tc_persistent(c.ks <= 3 ? 8 : 16)) {
        const int ntiles = (int)n;                                        // one pair tile per image
        const unsigned pairs = tc_pgrid(ntiles * 2) / 2;                 // one cluster per SM pair, at most one per tile
        k_conv_tc_2sm<<<2 * pairs, (2 + P2_NE) * 32, P2_SMEM_BYTES, st>>>(ts->act_map_h[a.in_planes][0], ts->act_map_h[a.in_planes][1],
                                                                           s.map_hi, s.map_lo, t, ntiles);
    } else if (tc_persistent(c.ks <= 3 ? 1 : 2))
        k_conv_tc_p<<<tc_pgrid((int)n * 4), TC_THREADS, P_SMEM_BYTES, st>>>(ts->act_map_h[a.in_planes][0], ts->act_map_h[a.in_planes][1],
                                                                           s.map_hi, s.map_lo, t, (int)n * 4);
    else
        k_conv_tc<<<dim3((unsigned)n, 256 / BN), TC_THREADS, SMEM_BYTES, st>>>(ts->act_map[a.in_planes][0], ts->act_map[a.in_planes][1],
                                                                           s.map_hi, s.map_lo, t);
    BSW_LAUNCH_CHECK();
    return BSW_OK;
}

// Heads: reads the planes the last ResNet conv wrote (in_planes), writes mu / sigma.  BSW_TC_HEADS=0 keeps the float32 SIMT
// kernel (A/B runs).
bool bsw_conv_tc_head_available(bsw_model *m, const ConvSlot &c) {
    static const bool off = getenv("BSW_TC_HEADS") && getenv("BSW_TC_HEADS")[0] == '0';
    return !off && m->tc_ready && c.tc_index >= 0 && c.Cout <= HD_N && c.Cin == m->d.reswidth;
}
int bsw_conv_tc_head(bsw_model *m, const ConvSlot &c, const ConvArgs &a, int64_t n, cudaStream_t st) {
    TcState *ts = (TcState *)m->tc_state;
    BSW_REQUIRE(ts && c.tc_index >= 0 && a.n_mu + a.n_sc <= HD_N, "bsw_conv_tc_head: conv has no tensor-core head weights");
    const TcSlot &s = (*(std::vector<TcSlot> *)m->tc_slots)[c.tc_index];
    HeadArgs h;
    h.taps = c.ks * c.ks; h.ks = c.ks; h.bias = c.bias; h.mu = a.mu; h.scale = a.scale; h.n_mu = a.n_mu; h.n_sc = a.n_sc;
    h.scale_kind = a.scale_kind; h.out_mode = a.out_mode; h.out_dim = a.out_dim;
    k_conv_tc_head<<<(unsigned)n, TC_THREADS, HD_SMEM_BYTES, st>>>(ts->act_map[a.in_planes][0], ts->act_map[a.in_planes][1], s.map_hi, s.map_lo, h);
    BSW_LAUNCH_CHECK();
    return BSW_OK;
}

// In-conv on the tensor cores: `given` (flat CHW float32) -> 32-channel bf16 planes -> k_conv_tc with one channel chunk
// per tap.  Epilogue as for every conv: T = ELU(raw) (the trunk), A planes = ELU(T) for the first ResNet layer.
int bsw_conv_tc_in(bsw_model *m, const ConvSlot &c, const ConvArgs &a, int64_t n, cudaStream_t st, int *launches) {
    TcState *ts = (TcState *)m->tc_state;
    BSW_REQUIRE(ts && c.tc_index >= 0, "bsw_conv_tc_in: conv has no tensor-core weights");
    const TcSlot &s = (*(std::vector<TcSlot> *)m->tc_slots)[c.tc_index];
    int64_t cnt = n * 256 * 32;
    k_given_to_planes<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(a.in, a.in_dim, c.Cin, a.in_mode, ts->inp[0], ts->inp[1], n);
    BSW_LAUNCH_CHECK();
    TcArgs t;
    t.taps = c.ks * c.ks; t.ks = c.ks;
    t.cchunks = 1;
    t.bias = c.bias; t.resid = nullptr; t.T = a.T; t.T_elu = a.T_elu;
    t.A_hi = a.A_planes >= 0 ? ts->act[a.A_planes][0] : nullptr;
    t.A_lo = a.A_planes >= 0 ? ts->act[a.A_planes][1] : nullptr;
    t.A_elu = a.A_elu;
    // in-convs: nine or twenty-five tiny k-blocks, all epilogue -> half-image tiles, two CTAs per SM (0.26 ms per launch at
    // 1024 images against 0.55 ms on the full tile)
    if (tc_persistent(4))
        k_conv_tc_p<<<tc_pgrid((int)n * 4), TC_THREADS, P_SMEM_BYTES, st>>>(ts->inp_map_h[0], ts->inp_map_h[1], s.map_hi, s.map_lo, t, (int)n * 4);
    else
        k_conv_tc_h<<<dim3((unsigned)n, 256 / BN, 2), TC_THREADS, H_SMEM_BYTES, st>>>(ts->inp_map_h[0], ts->inp_map_h[1], s.map_hi, s.map_lo, t);
    BSW_LAUNCH_CHECK();
    *launches += 2;
    return BSW_OK;
}
