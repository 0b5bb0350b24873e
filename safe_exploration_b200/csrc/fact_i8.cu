// The dense GEMMs of the setup factorisation (trailing updates of potrf, the block products of trtri) on the
// 5th-generation tensor cores: float64-grade products from int8 digit planes, like the variance contraction
// (tri_i8.cu) but with EIGHT balanced base-256 digits per operand (2^-63 of the row scale: every element keeps its
// float64 accuracy down to 2^-10 of its row's maximum) and the 36 digit-plane products a + c < 8, int32 accumulation
// in TMEM, one accumulator per diagonal a + c, float64 Horner recombination in the epilogue.  (Seven digits, 28
// products, measured 10-30x the float64 path's error on the predictive mean of ill-conditioned models: the error of a
// fixed-point row is relative to the row's MAXIMUM, and rows of a Cholesky factor span many orders of magnitude.)
//
//   fd_split    an operand block of a float64 matrix (optionally transposed / triangular) -> row scales (exact max-abs
//               over the k range) + 8 digit planes per (128-row tile, 64-column k-block), each plane an 8 KB
//               SWIZZLE_64B shared-memory image that the tensor core reads as it lies
//   gemm_i8d   C (M x N) = alpha A B^T + beta C from two plane sets; 128 x 64 output tile per CTA, bulk-copy ring of
//               two 96 KB stages, 20 MMAs per 32-deep k-step (two B planes per N = 128 instruction where the pair
//               exists), all 512 TMEM columns; flags as the DMMA gemm64 of setup.cu (triangular k trimming, lower tiles)
//
// Replaces the float64 DMMA GEMMs (mma.sync m8n8k4) of setup.cu where the k range is long enough to pay for the
// split: the K = 256 trailing update of the two-level potrf and the levels s >= 256 of trtri.  The reference
// interface behind it is GPy's posterior (Cholesky of K + noise I, woodbury_inv), ssm_gpy/gaussian_process.py:238-263.
#include <math.h>

#include <algorithm>

#include "segp_internal.cuh"
#include "tc_i8.cuh"

namespace segp {

constexpr int FD_S = 8;                          // digits per operand
constexpr int FD_TILE_M = 128;                   // rows of an A tile image (= TILE)
constexpr int FD_TILE_N = 64;                    // output columns per CTA: half of a tile image
constexpr int FD_A_TILE = FD_TILE_M * I8_KB;     // 8192 B, one digit plane of one (row tile, k-block)
constexpr int FD_B_HALF = FD_TILE_N * I8_KB;     // 4096 B
constexpr long FD_BLOCK = (long)FD_S * FD_A_TILE;   // 65536 B: the 8 planes of one (row tile, k-block)
constexpr int FD_STAGES = 2;
constexpr int FD_STAGE_BYTES = FD_S * (FD_A_TILE + FD_B_HALF);   // 98304
constexpr int FD_THREADS = 192;                  // producer, MMA issuer, 4 epilogue warps
constexpr size_t FD_SMEM = (size_t)FD_STAGES * FD_STAGE_BYTES + 1024 /* alignment slack */ + 128;
constexpr double FD_UNIT = 127.0 * 72057594037927936.0;   // 127 * 2^56: r in [-1, 1] -> integer of 63 bits + sign

// r in [-1, 1] -> V = rn(r 127 2^56) = d0 2^56 + d1 2^48 + ... + d7, d0 in [-127, 127], d1..d7 in [-128, 127].  The
// product r * FD_UNIT is rounded to 53 bits first, i.e. every ELEMENT keeps float64 relative accuracy as long as it is
// within 2^-10 of its row's maximum, and 2^-63 of the row maximum below that.
__device__ __forceinline__ void split_digits_fd(double r, int (&dg)[FD_S]) {
    long long v = __double2ll_rn(r * FD_UNIT);
#pragma unroll
    for (int a = FD_S - 1; a >= 1; --a) {
        const int d = (int)((v + 128) & 255) - 128;
        dg[a] = d;
        v = (v - d) >> 8;
    }
    dg[0] = (int)v;
}

// ------------------------------------------------------------------------------------------- operand view
// Operand element O[r][k] of batch entry z:  transposed ? X[k * ld + r] : X[r * ld + k], X advanced by z * zstride;
// tri 1: zero for k > r (lower-triangular operand), tri 2: zero for k < r (transpose of a lower-triangular block).
// Ragged last batch entry: lim = min(nominal, lim_total - z * zrows) bounds the rows (clip & 1) and/or the k range
// (clip & 2) -- same convention as GemmArgs::m_total / zrows.
struct FdSplitArgs {
    const double* x;
    long ld, zstride;
    int rows, cols;          // nominal operand size: rows % 128 == 0, cols % 64 == 0
    int transposed, tri, clip;
    int lim_total, zrows;
    unsigned long long* scale_bits;   // [batch][rows] max-abs as the bit pattern of a non-negative double
    int8_t* planes;                   // [batch][rows/128][cols/64][8][8192]
};

__device__ __forceinline__ bool fd_block_range(const FdSplitArgs& a, int z, int& rows_z, int& cols_z) {
    int lim = a.clip != 0 ? min(max(a.rows, a.cols), a.lim_total - z * a.zrows) : 0;
    rows_z = (a.clip & 1) ? min(a.rows, lim) : a.rows;
    cols_z = (a.clip & 2) ? min(a.cols, lim) : a.cols;
    return (int)blockIdx.y * FD_TILE_M < rows_z && (int)blockIdx.x * I8_KB < cols_z;
}
// 128 operand rows x 16 k of the block at (row tile blockIdx.y, k-block blockIdx.x), k chunk c, into shared memory
// with coalesced global reads whichever way the operand lies in memory: consecutive threads read consecutive
// addresses (along k for a row-major operand: 16 threads per 128-byte row segment; along r for a transposed one).
constexpr int FD_PITCH = 17;
__device__ __forceinline__ void fd_load_chunk(const FdSplitArgs& a, const double* __restrict__ x, int c,
                                              double (*tile)[FD_PITCH]) {
    const int r0 = blockIdx.y * FD_TILE_M, k0 = blockIdx.x * I8_KB + c * 16;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int e = i * FD_TILE_M + threadIdx.x;
        const int rl = a.transposed ? (e & (FD_TILE_M - 1)) : (e >> 4);
        const int kl = a.transposed ? (e >> 7) : (e & 15);
        const int r = r0 + rl, k = k0 + kl;
        double v = 0.0;
        if (!((a.tri == 1 && k > r) || (a.tri == 2 && k < r))) v = a.transposed ? x[(long)k * a.ld + r] : x[(long)r * a.ld + k];
        tile[rl][kl] = v;
    }
}

// grid (cols/64, rows/128, batch), 128 threads: thread = operand row
__global__ void __launch_bounds__(FD_TILE_M) fd_absmax_kernel(const FdSplitArgs a) {
    __shared__ double tile[FD_TILE_M][FD_PITCH];
    int rows_z, cols_z;
    const int z = blockIdx.z;
    if (!fd_block_range(a, z, rows_z, cols_z)) return;
    const double* __restrict__ x = a.x + (long)z * a.zstride;
    const int r = blockIdx.y * FD_TILE_M + threadIdx.x;
    double m = 0.0;
    for (int c = 0; c < I8_KB / 16; ++c) {
        __syncthreads();
        fd_load_chunk(a, x, c, tile);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) m = fmax(m, fabs(tile[threadIdx.x][j]));
    }
    if (m > 0.0) atomicMax(a.scale_bits + (long)z * a.rows + r, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(FD_TILE_M) fd_split_kernel(const FdSplitArgs a) {
    __shared__ double tile[FD_TILE_M][FD_PITCH];
    int rows_z, cols_z;
    const int z = blockIdx.z;
    if (!fd_block_range(a, z, rows_z, cols_z)) return;
    const double* __restrict__ x = a.x + (long)z * a.zstride;
    const int rl = threadIdx.x;
    const int r = blockIdx.y * FD_TILE_M + rl;
    const double sc = __longlong_as_double((long long)a.scale_bits[(long)z * a.rows + r]);
    const double inv = sc > 0.0 ? 1.0 / sc : 0.0;
    const int nkb = a.cols / I8_KB;
    int8_t* dst = a.planes + (((long)z * (a.rows / FD_TILE_M) + blockIdx.y) * nkb + blockIdx.x) * FD_BLOCK;
#pragma unroll 1
    for (int c = 0; c < I8_KB / 16; ++c) {
        __syncthreads();
        fd_load_chunk(a, x, c, tile);
        __syncthreads();
        uint32_t pk[FD_S][4];
#pragma unroll
        for (int p = 0; p < FD_S; ++p) pk[p][0] = pk[p][1] = pk[p][2] = pk[p][3] = 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int dg[FD_S];
            // |v| <= sc, but v * (1 / sc) may exceed 1 by an ulp: clamp so the leading digit stays within +-127
            const double q = fmax(-1.0, fmin(1.0, tile[rl][j] * inv));
            split_digits_fd(q, dg);
#pragma unroll
            for (int p = 0; p < FD_S; ++p) pk[p][j >> 2] |= (uint32_t)(dg[p] & 0xff) << ((j & 3) * 8);
        }
        const int off = sw64_offset(rl, c * 16);
#pragma unroll
        for (int p = 0; p < FD_S; ++p)
            *reinterpret_cast<uint4*>(dst + (long)p * FD_A_TILE + off) = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
    }
}

size_t fd_plane_bytes(int rows, int cols, int batch) {
    return (size_t)batch * (rows / FD_TILE_M) * (cols / I8_KB) * FD_BLOCK;
}

// scales (double[batch * rows], written as bit patterns) and planes of one operand; three launches
int fd_split(const double* x, long ld, long zstride, int rows, int cols, int transposed, int tri, int clip, int lim_total,
             int zrows, int batch, double* scale, int8_t* planes, cudaStream_t st) {
    if (rows <= 0 || cols <= 0 || batch <= 0) return SEGP_OK;
    if (rows % FD_TILE_M != 0 || cols % I8_KB != 0) {
        set_error("fd_split: operand of %d x %d is not a multiple of 128 x 64", rows, cols);
        return SEGP_ERR_INVALID;
    }
    FdSplitArgs a{};
    a.x = x;
    a.ld = ld;
    a.zstride = zstride;
    a.rows = rows;
    a.cols = cols;
    a.transposed = transposed;
    a.tri = tri;
    a.clip = clip;
    a.lim_total = lim_total;
    a.zrows = zrows;
    a.scale_bits = reinterpret_cast<unsigned long long*>(scale);
    a.planes = planes;
    SEGP_CUDA_CHECK(cudaMemsetAsync(scale, 0, (size_t)batch * rows * sizeof(double), st));
    dim3 grid((unsigned)(cols / I8_KB), (unsigned)(rows / FD_TILE_M), (unsigned)batch);
    fd_absmax_kernel<<<grid, FD_TILE_M, 0, st>>>(a);
    fd_split_kernel<<<grid, FD_TILE_M, 0, st>>>(a);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// ------------------------------------------------------------------------------------------- gemm_i8d
// All MMAs of one k-block (two k-steps of 32): W-style plane loop of tri_i8m -- A plane a against B planes 0..7-a,
// accumulator slot a + c; B planes c and c + 1 are adjacent 4 KB images and their products belong to adjacent slots,
// so ONE N = 128 instruction does both.  20 MMAs per k-step (16 x N = 128, 4 x N = 64).
__device__ __forceinline__ void fd_issue(uint32_t tmem_base, uint32_t sa, bool first) {
    constexpr uint32_t idesc1 = make_i8_idesc(FD_TILE_M, FD_TILE_N);
    constexpr uint32_t idesc2 = make_i8_idesc(FD_TILE_M, 2 * FD_TILE_N);
    const uint32_t sb = sa + (uint32_t)(FD_S * FD_A_TILE);
#pragma unroll
    for (int ks = 0; ks < I8_KB / 32; ++ks) {
#pragma unroll
        for (int pa = 0; pa < FD_S; ++pa) {
            const uint64_t adesc = make_sw64_desc(sa + pa * FD_A_TILE + ks * 32);
            const uint32_t acc = (uint32_t)(!first || ks != 0 || pa != 0);   // plane 0 touches every slot first
#pragma unroll
            for (int pc = 0; pc < FD_S - pa; pc += 2) {
                const uint64_t bdesc = make_sw64_desc(sb + pc * FD_B_HALF + ks * 32);
                const bool two = pc + 1 < FD_S - pa;
                tc_mma_i8(tmem_base + (uint32_t)((pa + pc) * FD_TILE_N), adesc, bdesc, two ? idesc2 : idesc1, acc);
            }
        }
    }
}

__global__ void __launch_bounds__(FD_THREADS, 1) gemm_i8d_kernel(const GemmI8Args g) {
    const int z = blockIdx.z;
    const int m_z = min(g.m, g.m_total - z * g.zrows);
    const int row0 = blockIdx.y * FD_TILE_M;
    const int col0 = blockIdx.x * FD_TILE_N;
    if (row0 >= m_z) return;
    if ((g.flags & GEMM_C_LOWER) && col0 > row0 + FD_TILE_M - 1) return;
    int kb0 = 0, kb1 = g.k / I8_KB;
    if (g.flags & GEMM_A_LOWER) kb1 = min(kb1, (row0 + FD_TILE_M) / I8_KB);
    if (g.flags & GEMM_B_LOWER) kb0 = col0 / I8_KB;
    const int nk = kb1 - kb0;
    double* __restrict__ C = g.c + (long)z * g.z_c;
    const double* __restrict__ as = g.as + (long)z * g.z_as;
    const double* __restrict__ bs = g.bs + (long)z * g.z_bs;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (nk <= 0) {   // empty k range (block-uniform): C = beta C
        if (warp >= 2) {
            const int row = row0 + (warp & 3) * 32 + lane;
            for (int j = 0; j < FD_TILE_N; ++j) {
                double* dst = C + (long)row * g.ldc + col0 + j;
                *dst = g.beta != 0.0 ? g.beta * *dst : 0.0;
            }
        }
        return;
    }

    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_addr(smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;
    unsigned char* tail = smem_raw + (stage0 - raw) + (size_t)FD_STAGES * FD_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * FD_STAGES + 1);
    const uint32_t bar0 = smem_addr(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (FD_STAGES + s); };
    const uint32_t tmem_full_bar = bar0 + 8u * (2 * FD_STAGES);

    if (threadIdx.x == 0) {
        for (int s = 0; s < FD_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int8_t* asrc = g.ap + (long)z * g.z_ap + ((long)blockIdx.y * g.a_kb) * FD_BLOCK;
            const int8_t* bsrc = g.bp + (long)z * g.z_bp + ((long)(blockIdx.x >> 1) * g.b_kb) * FD_BLOCK +
                                 (long)(blockIdx.x & 1) * FD_B_HALF;
            for (int it = 0; it < nk; ++it) {
                const int s = it % FD_STAGES;
                if (it >= FD_STAGES) mbar_wait(empty_bar(s), (uint32_t)((it / FD_STAGES - 1) & 1));
                const uint32_t dst = stage0 + (uint32_t)s * FD_STAGE_BYTES;
                const long kb = kb0 + it;
                mbar_expect_tx(full_bar(s), FD_STAGE_BYTES);
                bulk_g2s(dst, asrc + kb * FD_BLOCK, (uint32_t)FD_BLOCK, full_bar(s));
#pragma unroll
                for (int p = 0; p < FD_S; ++p)
                    bulk_g2s(dst + (uint32_t)(FD_S * FD_A_TILE + p * FD_B_HALF), bsrc + kb * FD_BLOCK + (long)p * FD_A_TILE,
                             FD_B_HALF, full_bar(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it % FD_STAGES;
                mbar_wait(full_bar(s), (uint32_t)((it / FD_STAGES) & 1));
                tc_fence_after();
                fd_issue(tmem_base, stage0 + (uint32_t)s * FD_STAGE_BYTES, it == 0);
                tc_commit(empty_bar(s));
            }
            tc_commit(tmem_full_bar);
        }
    } else {
        // epilogue: thread = output row (TMEM lane), two chunks of 32 columns; Horner over the 8 diagonals in float64
        // (the first steps are exact, the last ones round at 2^-53 of the running value).  The scaled 32 x 32 block
        // then crosses a shared-memory transpose (the drained pipeline stages) so that C is read and written with the
        // lanes along a row: 256-byte segments instead of 32 rows x 16 bytes per instruction -- a read-modify-write
        // of C at the row stride ran the K = 256 trailing update 19x slower (2.4 ms instead of 0.13 ms at C4).
        const int q = warp & 3;
        const int row = row0 + q * 32 + lane;
        // sum_g C_g 256^(14-g) = 256^7 * Horner value; operand values are V / FD_UNIT
        const double f = g.alpha * as[row] * (72057594037927936.0 / FD_UNIT / FD_UNIT);
        if (g.beta != 0.0) {
            // the read-modify-write of C: pull this warp's 32 x 64 block (32 rows x 4 lines) into L2 while the MMAs run
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double* pf = C + (long)(row0 + q * 32 + (i * 32 + lane) / 4) * g.ldc + col0 + ((i * 32 + lane) & 3) * 16;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            }
        }
        mbar_wait(tmem_full_bar, 0u);
        tc_fence_after();
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        double* xpose = reinterpret_cast<double*>(smem_raw + (stage0 - raw)) + (warp - 2) * (32 * 33);
#pragma unroll 1
        for (int chunk = 0; chunk < FD_TILE_N / 32; ++chunk) {
            uint32_t v[32];
            double acc[32];
            tmem_ld32(tq + (uint32_t)(chunk * 32), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = i8_cvt_s32(v[j]);
#pragma unroll
            for (int s = 1; s < FD_S; ++s) {
                tmem_ld32(tq + (uint32_t)(s * FD_TILE_N + chunk * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = fma(acc[j], I8_BASE, i8_cvt_s32(v[j]));
            }
            const double bsc = bs[col0 + chunk * 32 + lane];   // scale of column `lane` of the chunk
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) xpose[lane * 33 + j] = acc[j] * f;
            __syncwarp();
            double* dst = C + (long)(row0 + q * 32) * g.ldc + col0 + chunk * 32 + lane;
            double old[32];   // all 32 row reads in flight before the first store
            if (g.beta != 0.0) {
#pragma unroll
                for (int r = 0; r < 32; ++r) old[r] = dst[(long)r * g.ldc];
            }
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                double out = xpose[r * 33 + lane] * bsc;
                if (g.beta != 0.0) out = fma(g.beta, old[r], out);
                dst[(long)r * g.ldc] = out;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

int launch_gemm_i8d(const GemmI8Args& g, int batch, cudaStream_t st) {
    if (g.m <= 0 || g.n <= 0 || batch <= 0) return SEGP_OK;
    // the B planes come in 128-row tiles (a CTA takes one 64-row half); int32 accumulators: 8 pairs x k x 128^2 < 2^31
    if (g.m % FD_TILE_M != 0 || g.n % FD_TILE_M != 0 || g.k % I8_KB != 0 || g.k > I8_MAX_NPAD / 2) {
        set_error("gemm_i8d: %d x %d x %d is not a multiple of 128 x 128 x 64 (or k beyond %ld)", g.m, g.n, g.k,
                  I8_MAX_NPAD / 2);
        return SEGP_ERR_INVALID;
    }
    static bool attr_set[64] = {};
    if (first_call_on_device(attr_set))
        SEGP_CUDA_CHECK(cudaFuncSetAttribute(gemm_i8d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FD_SMEM));
    dim3 grid((unsigned)(g.n / FD_TILE_N), (unsigned)(g.m / FD_TILE_M), (unsigned)batch);
    gemm_i8d_kernel<<<grid, FD_THREADS, FD_SMEM, st>>>(g);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// ------------------------------------------------------------------------------------------- scratch
size_t fd_scratch_plane_bytes(int n_pad) {
    // the largest operand the factorisation splits: max over the trtri levels s >= 256 of pairs * s * s, and the
    // potrf panel (n_pad x 256)
    size_t elems = (size_t)n_pad * 256;
    for (long s = 256; s < n_pad; s *= 2) {
        const long pairs = (n_pad - s + 2 * s - 1) / (2 * s);
        elems = std::max(elems, (size_t)(pairs * s * s));
    }
    return elems * FD_S;
}

// ------------------------------------------------------------------------------------------- self-test
// C = alpha A op(B) + beta C through split + tensor-core GEMM, host pointers; A is m x k; B is n x k (trans_b) or
// k x n; tri flags as GEMM_*_LOWER.
int gemm_i8d_selftest(int m, int n, int k, const double* h_a, const double* h_b, double* h_c, double alpha, double beta,
                       int trans_b, int flags) {
    if (m % FD_TILE_M != 0 || n % FD_TILE_M != 0 || k % I8_KB != 0) {
        set_error("gemm_i8d_selftest: m, n multiples of 128 and k a multiple of 64 required");
        return SEGP_ERR_INVALID;
    }
    double *a = nullptr, *b = nullptr, *c = nullptr, *as = nullptr, *bs = nullptr;
    int8_t *ap = nullptr, *bp = nullptr;
    int rc = SEGP_OK;
    auto dev_alloc = [](auto** p, size_t count) -> int {
        if (cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(**p)) != cudaSuccess) {
            set_error("gemm_i8d_selftest: out of device memory");
            return SEGP_ERR_CUDA;
        }
        return SEGP_OK;
    };
    auto dev_free = [](auto* p) {
        if (p != nullptr) cudaFree(p);
    };
    do {
        if ((rc = dev_alloc(&a, (size_t)m * k)) != SEGP_OK) break;
        if ((rc = dev_alloc(&b, (size_t)n * k)) != SEGP_OK) break;
        if ((rc = dev_alloc(&c, (size_t)m * n)) != SEGP_OK) break;
        if ((rc = dev_alloc(&as, (size_t)m)) != SEGP_OK) break;
        if ((rc = dev_alloc(&bs, (size_t)n)) != SEGP_OK) break;
        if ((rc = dev_alloc(&ap, fd_plane_bytes(m, k, 1))) != SEGP_OK) break;
        if ((rc = dev_alloc(&bp, fd_plane_bytes(n, k, 1))) != SEGP_OK) break;
        cudaMemcpy(a, h_a, (size_t)m * k * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(b, h_b, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(c, h_c, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice);
        if ((rc = fd_split(a, k, 0, m, k, 0, (flags & GEMM_A_LOWER) ? 1 : 0, 0, 0, 0, 1, as, ap, nullptr)) != SEGP_OK) break;
        if (trans_b)
            rc = fd_split(b, k, 0, n, k, 0, 0, 0, 0, 0, 1, bs, bp, nullptr);
        else
            rc = fd_split(b, n, 0, n, k, 1, (flags & GEMM_B_LOWER) ? 2 : 0, 0, 0, 0, 1, bs, bp, nullptr);
        if (rc != SEGP_OK) break;
        GemmI8Args g{};
        g.ap = ap;
        g.as = as;
        g.a_kb = k / I8_KB;
        g.bp = bp;
        g.bs = bs;
        g.b_kb = k / I8_KB;
        g.c = c;
        g.ldc = n;
        g.m = m;
        g.n = n;
        g.k = k;
        g.alpha = alpha;
        g.beta = beta;
        g.flags = flags;
        g.m_total = m;
        g.zrows = 0;
        if ((rc = launch_gemm_i8d(g, 1, nullptr)) != SEGP_OK) break;
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            set_error("gemm_i8d_selftest: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
            break;
        }
        cudaMemcpy(h_c, c, (size_t)m * n * sizeof(double), cudaMemcpyDeviceToHost);
    } while (0);
    dev_free(a);
    dev_free(b);
    dev_free(c);
    dev_free(as);
    dev_free(bs);
    dev_free(ap);
    dev_free(bp);
    return rc;
}

}  // namespace segp
