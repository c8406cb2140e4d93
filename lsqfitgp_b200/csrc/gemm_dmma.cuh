// FP64 GEMM on the Blackwell FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
//   C[i][j] = (beta0 ? 0 : C[i][j]) + alpha * sum_k Aop[i][k] * Bop[j][k]
//
// Operand storage (all row-major with a leading dimension in doubles):
//   a_kmajor : Aop[i][k] = A[i*lda + k]      (k contiguous)
//   !a_kmajor: Aop[i][k] = A[k*lda + i]      (i contiguous, "transposed")
//   b_kmajor : Bop[j][k] = B[j*ldb + k]
//   !b_kmajor: Bop[j][k] = B[k*ldb + j]
//
// This single kernel family is the trailing-update (SYRK), the TRSM (as multiplication by
// the inverted 128x128 diagonal block) and the TRTRI/LAUUM building block of the blocked
// Cholesky that replaces jax.scipy.linalg.cholesky / solve_triangular in the reference
// (reference: src/lsqfitgp/_linalg/_decomp.py:388,402-403,467-472).
//
// Two tile configurations: 128x128x16 with 16 warps (4x4, warp tile 32x32) for large problems, and
// 64x64x16 with 4 warps (2x2) for the small, latency-bound products on the panel critical path.
// Multi-stage cp.async pipeline.  tcgen05/TMEM have no f64 kind, so the warp-level DMMA is the FP64
// tensor path on sm_100a.
#pragma once
#include "common.cuh"

namespace lgp {

enum GemmFlags : int {
    GEMM_LOWER = 1,       // C is square; compute/store only j <= i
    GEMM_BETA0 = 2,       // overwrite C instead of accumulating
    GEMM_A_LOWER_K = 4,   // Aop[i][k] == 0 for k > i  (skip k-tiles beyond the row tile)
    GEMM_B_LOWER_K = 8,   // Bop[j][k] == 0 for k > j
    GEMM_A_UPPER_K = 16,  // Aop[i][k] == 0 for k < i
    GEMM_B_UPPER_K = 32,  // Bop[j][k] == 0 for k < j
    GEMM_INPLACE_A = 64,   // C aliases A (rows of C depend on the same rows of A only): needs N <= BN
    GEMM_INPLACE_B = 128,  // C aliases B: needs M <= BM
    GEMM_FORCE_BIG = 256,  // always use the 128x128 configuration
};

// Extra destinations of the epilogue (fused compute -> broadcast over peer memory): every C entry is also stored at
// dst[i] + row*ld + col.  dst[i] are device pointers valid on THIS GPU: local memory, peer memory mapped over
// NVLink (unicast, one store per peer), or ONE NVSwitch multicast address (`multimem` != 0: a single multimem.st
// reaches every GPU of the group).  Used by the panel TRSM of the block-cyclic Cholesky (dist.cu).
constexpr int GEMM_MAX_MIRRORS = 8;
struct GemmMirror {
    int n;         // number of destinations (0: none)
    int multimem;  // destinations are multicast addresses
    int64_t ld;
    double *dst[GEMM_MAX_MIRRORS];
};

struct GemmParams {
    const double *A;
    const double *B;
    double *C;
    int M, N, K;
    int64_t lda, ldb, ldc;
    double alpha;
    int flags;
    int tiles_m, tiles_n;
    const double *scale;  // optional: C[i][j] = (alpha acc[i][j] * scale[j]) * scale[i] (+ C): the equilibration scales of the
                          // inverse applied in the epilogue instead of a separate pass over the matrix; needs N entries
    GemmMirror mir;
};

__device__ __forceinline__ void gemm_mirror_store(const GemmMirror &m, int64_t off, double v) {
    if (m.multimem) {
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(m.dst[0] + off), "d"(v) : "memory");
    } else {
        for (int i = 0; i < m.n; i++) m.dst[i][off] = v;
    }
}
// two adjacent entries, 16-byte aligned: one 128-bit store (multimem.st has no .v2.f64 form; the bit pattern goes
// through the .v4.f32 form, which is the same STG.128 on the multicast address)
__device__ __forceinline__ void gemm_mirror_store2(const GemmMirror &m, int64_t off, double v0, double v1) {
    if (m.multimem) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(m.dst[0] + off),
                     "f"(__int_as_float(__double2loint(v0))), "f"(__int_as_float(__double2hiint(v0))),
                     "f"(__int_as_float(__double2loint(v1))), "f"(__int_as_float(__double2hiint(v1)))
                     : "memory");
    } else {
        for (int i = 0; i < m.n; i++) *reinterpret_cast<double2 *>(m.dst[i] + off) = make_double2(v0, v1);
    }
}

constexpr int GEMM_BK = 16;

template <int BM_, int BN_, int WARPS_M_, int WARPS_N_, int STAGES_, int MINBLOCKS_ = 1>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, STAGES = STAGES_;
    static constexpr int MINBLOCKS = MINBLOCKS_;
    static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
    static constexpr int MI = BM / (8 * WARPS_M);  // 8-row DMMA fragments per warp (rows)
    static constexpr int NJ = BN / (8 * WARPS_N);  // 8-col DMMA fragments per warp (cols)
    // k-major tile: rows x 128 B, 16-B chunks XOR-swizzled by ((row&1)<<2): conflict-free LDS.128
    // m-major tile: 16 k-rows x (rows+2) doubles: conflict-free LDS.64 for the k = 8g+2q+t slot order
    static constexpr int A_BYTES = (BM + 2) * GEMM_BK * 8;
    static constexpr int B_BYTES = (BN + 2) * GEMM_BK * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES;
};

// Main configuration: 64x64 CTA tile, 4 warps (2x2, warp tile 32x32), 3 CTAs per SM.  Several small,
// mutually unsynchronised CTAs per SM keep the DMMA pipe fed across each other's per-k-tile barriers
// (measured on B200: 32.9 TFLOP/s NT 8192^3 and 31.4 TFLOP/s on the rank-512 trailing SYRK, against 30.8 / 27.3
// for one 128x128 16-warp CTA per SM; FP64 DMMA peak 37.0).
using GemmBig = GemmCfg<64, 64, 2, 2, 4, 3>;
using GemmTall = GemmCfg<64, 128, 2, 4, 4, 2>;  // C aliases A (right-TRSM leaf): all N <= 128 columns in one tile
using GemmWide = GemmCfg<128, 64, 4, 2, 4, 2>;  // C aliases B (left-solve leaf): all M <= 128 rows in one tile
// Latency configurations for products whose grid would not even fill the SMs once (the panel chain of the factorisation's
// tail, small matrices): a quarter of the tile area, so 4x the CTAs and a quarter of the dependent k-loop time per CTA
// (a 64x64x128 tile is 8 k-tiles x 1024 cycles of DMMA issue on ONE SM = 4.2 us; 32x32: 1 us).  Small footprints (35 /
// 57 KB, 128 threads) also fit into the slot a finishing GemmBig CTA frees.
using GemmSmall = GemmCfg<32, 32, 2, 2, 4, 4>;
using GemmTallSmall = GemmCfg<16, 128, 1, 4, 3, 3>;
using GemmWideSmall = GemmCfg<128, 16, 4, 1, 3, 3>;

template <bool KMAJ, int ROWS, int THREADS>
__device__ __forceinline__ void gemm_load_tile(uint32_t smem_tile, const double *__restrict__ G, int64_t ld,
                                               int r0, int rows_total, int k0, int k_total, int tid) {
    constexpr int CHUNKS = ROWS * GEMM_BK / 2;
    static_assert(CHUNKS % THREADS == 0, "tile chunks must divide evenly among threads");
    constexpr int ITERS = CHUNKS / THREADS;
    if (KMAJ) {
#pragma unroll
        for (int it = 0; it < ITERS; it++) {
            int id = tid + it * THREADS;
            int row = id >> 3, c = id & 7;
            int gr = r0 + row, gk = k0 + 2 * c;
            int bytes = 0;
            if (gr < rows_total) {
                int rem = (k_total - gk) * 8;
                bytes = rem >= 16 ? 16 : (rem > 0 ? rem : 0);
            }
            const double *src = bytes ? (G + (int64_t)gr * ld + gk) : G;
            uint32_t dst = smem_tile + row * 128 + ((c ^ ((row & 1) << 2)) << 4);
            cp_async16(dst, src, bytes);
        }
    } else {
        constexpr int CPR = ROWS / 2;  // 16-byte chunks per k-row
#pragma unroll
        for (int it = 0; it < ITERS; it++) {
            int id = tid + it * THREADS;
            int krow = id / CPR, c = id % CPR;
            int gk = k0 + krow, gr = r0 + 2 * c;
            int bytes = 0;
            if (gk < k_total) {
                int rem = (rows_total - gr) * 8;
                bytes = rem >= 16 ? 16 : (rem > 0 ? rem : 0);
            }
            const double *src = bytes ? (G + (int64_t)gk * ld + gr) : G;
            uint32_t dst = smem_tile + krow * ((ROWS + 2) * 8) + c * 16;
            cp_async16(dst, src, bytes);
        }
    }
}

// Loop-invariant form of gemm_load_tile for k-tiles that lie entirely inside [0, k_total): the source pointer of
// this thread's first chunk is computed once and advanced by one k-tile per call, the row/column validity is folded
// into a per-iteration byte count, and the shared-memory offsets of the ITERS chunks differ by compile-time constants.
// (The generic loader costs ~25 integer instructions and two branches per 16-byte copy, which starved the DMMA pipe.)
template <bool KMAJ, int ROWS, int THREADS>
struct TileLoader {
    static constexpr int CHUNKS = ROWS * GEMM_BK / 2;
    static constexpr int ITERS = CHUNKS / THREADS;
    static constexpr int ROWS_PER_IT = KMAJ ? THREADS / 8 : THREADS / (ROWS / 2);  // tile rows (k-major) / k-rows per iter
    static_assert(CHUNKS % THREADS == 0, "tile chunks must divide evenly among threads");
    static_assert(!KMAJ || (ROWS_PER_IT % 2 == 0), "swizzle parity must not change between iterations");
    const char *src;     // global address of chunk 0 of the next k-tile
    int64_t it_stride;   // bytes between the chunks of consecutive iterations
    int64_t kt_stride;   // bytes between consecutive k-tiles
    uint32_t dst;        // shared-memory offset of chunk 0 inside a stage tile
    uint32_t bytes;      // 4 bits per iteration: valid bytes / 4 (0, 2 or 4); a zero-size copy reads nothing, so
                         // the addresses of out-of-range rows are never dereferenced

    __device__ __forceinline__ void init(const double *__restrict__ G, int64_t ld, int r0, int rows_total, int k0,
                                         int tid) {
        bytes = 0;
        if (KMAJ) {
            const int row = tid >> 3, c = tid & 7;
            dst = row * 128 + ((c ^ ((row & 1) << 2)) << 4);
            int gr = r0 + row;
            src = reinterpret_cast<const char *>(G + (int64_t)min(gr, rows_total - 1) * ld + k0 + 2 * c);
            it_stride = (int64_t)ROWS_PER_IT * ld * 8;
            kt_stride = GEMM_BK * 8;
#pragma unroll
            for (int it = 0; it < ITERS; it++) {
                int g = gr + it * ROWS_PER_IT;
                bytes |= (g < rows_total ? 4u : 0u) << (4 * it);
            }
        } else {
            constexpr int CPR = ROWS / 2;
            const int krow = tid / CPR, c = tid % CPR;
            dst = krow * ((ROWS + 2) * 8) + c * 16;
            const int gr = r0 + 2 * c;
            const int rem = rows_total - gr;
            const uint32_t b = rem >= 2 ? 4u : (rem == 1 ? 2u : 0u);
            src = reinterpret_cast<const char *>(G + (int64_t)(k0 + krow) * ld + (b ? gr : 0));
            it_stride = (int64_t)ROWS_PER_IT * ld * 8;
            kt_stride = (int64_t)GEMM_BK * ld * 8;
#pragma unroll
            for (int it = 0; it < ITERS; it++) bytes |= b << (4 * it);
        }
    }
    __device__ __forceinline__ void load(uint32_t smem_tile) {
#pragma unroll
        for (int it = 0; it < ITERS; it++) {
            constexpr int DST_STRIDE = KMAJ ? ROWS_PER_IT * 128 : ROWS_PER_IT * ((ROWS + 2) * 8);
            cp_async16(smem_tile + dst + it * DST_STRIDE, src + it * it_stride, ((bytes >> (4 * it)) & 15u) << 2);
        }
        src += kt_stride;
    }
    __device__ __forceinline__ void skip() { src += kt_stride; }
};

// Fragment for MMA slot q (= lane%4) within 8-k group g: k = 8g + 2q + t, t in {0,1}.
template <bool KMAJ, int ROWS>
__device__ __forceinline__ void gemm_load_frag(const unsigned char *tile, int row, int g, int q, double &v0,
                                               double &v1) {
    if (KMAJ) {
        int c = (4 * g + q) ^ ((row & 1) << 2);
        double2 v = *reinterpret_cast<const double2 *>(tile + row * 128 + (c << 4));
        v0 = v.x;
        v1 = v.y;
    } else {
        int k = 8 * g + 2 * q;
        const double *p = reinterpret_cast<const double *>(tile) + k * (ROWS + 2) + row;
        v0 = p[0];
        v1 = p[ROWS + 2];
    }
}

// m-major tile ([k][row], row stride ROWS + 2 doubles): ONE 128-bit load gives the entries of two adjacent rows at slot
// k, which feed two different 8x8 MMA tiles.  The MMA does not care which matrix row sits in lane-row lr of a tile as
// long as the accumulators are written back with the same map, so in m-major operands MMA tile 2p holds rows
// 16p + 2 lr and tile 2p + 1 rows 16p + 2 lr + 1 (k-major operands: tile i holds rows 8i + lr).  Same LDS count as the
// k-major path (the former two 64-bit loads per fragment cost 14 % of the GEMM rate).  Conflict-free: the 8 lanes of
// a quarter-warp (lr in {0,1} x q) hit offsets 32 q + 16 lr (mod 128 B) because 2 (ROWS + 2) 8 = 32 (mod 128).
template <int ROWS>
__device__ __forceinline__ void gemm_load_frag_mpair(const unsigned char *tile, int row, int k, double &v_even,
                                                     double &v_odd) {
    static_assert((2 * (ROWS + 2) * 8) % 128 == 32, "m-major row stride must keep the paired loads conflict-free");
    const double2 v = *reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(tile) + k * (ROWS + 2) + row);
    v_even = v.x;
    v_odd = v.y;
}

// Grouped rasterisation of the lower-triangular tile enumeration (R == 1): bands of GROUP tile rows, inside a band
// column by column (rows max(tn, r0)..r1 of column tn), so that a wave of CTAs shares GROUP row strips of A and
// ~wave/GROUP column strips of B out of L2.  Row by row, a wave spans one or two tile rows and streams EVERY column strip
// of B: with K = 1024 that is 160 MB per tile row at n = 20 000, more than L2 (the trailing updates of one factorisation read
// 64 GB from DRAM that way, profiles/traffic_chol20k_r2.txt; 28 GB grouped, traffic_chol20k_r2c.txt).  Rows 0..r0-1 hold
// r0 (r0+1)/2 tiles, so the band of tile b is the band of its row `row` in the row-by-row order.
__host__ __device__ __forceinline__ void gemm_lower_grouped_tile(long long b, int row, int tiles_m, int &tm, int &tn) {
    constexpr int GROUP = 16;
    const int r0 = (row / GROUP) * GROUP;
    const int gs = (tiles_m - r0) < GROUP ? (tiles_m - r0) : GROUP;
    int local = (int)(b - (long long)r0 * (r0 + 1) / 2);
    if (local < r0 * gs) {
        tn = local / gs;
        tm = r0 + local % gs;
    } else {
        local -= r0 * gs;
        int c = 0;
        while (local >= gs - c) {
            local -= gs - c;
            c++;
        }
        tn = r0 + c;
        tm = r0 + c + local;
    }
}

template <class Cfg, bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINBLOCKS) gemm_dmma_kernel(const GemmParams p) {
    static_assert(Cfg::MI % 2 == 0 && Cfg::NJ % 2 == 0, "paired m-major fragments need an even number of MMA tiles");
    constexpr int BM = Cfg::BM, BN = Cfg::BN, MI = Cfg::MI, NJ = Cfg::NJ, STAGES = Cfg::STAGES;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
    const int lr = lane >> 2, q = lane & 3;

    int tm, tn;
    if (p.flags & GEMM_LOWER) {
        // blockIdx.x enumerates the tiles that touch the lower triangle: with BM = R*BN row tm has R*(tm+1) column
        // tiles, so rows 0..tm-1 hold R*tm*(tm+1)/2 tiles
        constexpr int R = BM >= BN ? BM / BN : 1;  // (BM < BN configurations are never launched with LOWER)
        long long b = blockIdx.x;
        tm = (int)((sqrt(8.0 * (double)b / R + 1.0) - 1.0) * 0.5);
        while ((long long)R * (tm + 1) * (tm + 2) / 2 <= b) tm++;
        while ((long long)R * tm * (tm + 1) / 2 > b) tm--;
        tn = (int)(b - (long long)R * tm * (tm + 1) / 2);
        if (R == 1) gemm_lower_grouped_tile(b, tm, p.tiles_m, tm, tn);
    } else {
        // Grouped rasterisation: the CTAs in flight at any time (one wave = 3 per SM) cover GROUP row tiles x ~wave/GROUP
        // column tiles, so that both operand streams are re-used out of L2.  (With the plain column-of-tiles-fastest
        // order a wave spans ~440 row tiles of ONE tile column: every CTA streams its own A strip and all of A is re-read
        // from HBM once per tile column: 62 GB of DRAM reads for an 8192^3 product, profiles/gemm_dmma_ncu_r1.txt.)
        constexpr int GROUP = 16;
        const long long b = blockIdx.x;
        const long long per_group = (long long)GROUP * p.tiles_n;
        const int first_m = (int)(b / per_group) * GROUP;
        const int gsz = min(GROUP, p.tiles_m - first_m);
        const int rem = (int)(b % per_group);
        tm = first_m + rem % gsz;
        tn = rem / gsz;
    }
    const int m0 = tm * BM, n0 = tn * BN;

    int k_begin = 0, k_end = p.K;
    if (p.flags & GEMM_A_LOWER_K) k_end = min(k_end, m0 + BM);
    if (p.flags & GEMM_B_LOWER_K) k_end = min(k_end, n0 + BN);
    if (p.flags & GEMM_A_UPPER_K) k_begin = max(k_begin, m0);
    if (p.flags & GEMM_B_UPPER_K) k_begin = max(k_begin, n0);
    k_begin &= ~(GEMM_BK - 1);
    const int KT = k_end > k_begin ? (k_end - k_begin + GEMM_BK - 1) / GEMM_BK : 0;

    const uint32_t smem_base = smem_u32(smem);

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // k-tiles [0, KT_full) lie entirely inside [k_begin, k_end): loop-invariant loaders; a partial last tile (K not
    // a multiple of 16) goes through the generic loader
    const int KT_full = (k_end - k_begin) / GEMM_BK < KT ? (k_end - k_begin) / GEMM_BK : KT;
    TileLoader<A_KMAJ, BM, Cfg::THREADS> la;
    TileLoader<B_KMAJ, BN, Cfg::THREADS> lb;
    la.init(p.A, p.lda, m0, p.M, k_begin, tid);
    lb.init(p.B, p.ldb, n0, p.N, k_begin, tid);
    auto load_stage = [&](int nk) {
        const uint32_t sa = smem_base + (nk % STAGES) * Cfg::STAGE_BYTES;
        if (nk < KT_full) {
            la.load(sa);
            lb.load(sa + Cfg::A_BYTES);
        } else if (nk < KT) {
            int k0 = k_begin + nk * GEMM_BK;
            gemm_load_tile<A_KMAJ, BM, Cfg::THREADS>(sa, p.A, p.lda, m0, p.M, k0, k_end, tid);
            gemm_load_tile<B_KMAJ, BN, Cfg::THREADS>(sa + Cfg::A_BYTES, p.B, p.ldb, n0, p.N, k0, k_end, tid);
        }
    };

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        load_stage(s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        load_stage(kt + STAGES - 1);
        cp_async_commit();
        const int s = kt % STAGES;
        const unsigned char *at = smem + s * Cfg::STAGE_BYTES;
        const unsigned char *bt = at + Cfg::A_BYTES;
#pragma unroll
        for (int g = 0; g < 2; g++) {
            double af[MI][2], bf[NJ][2];
            if (A_KMAJ) {
#pragma unroll
                for (int i = 0; i < MI; i++)
                    gemm_load_frag<true, BM>(at, wm * (8 * MI) + i * 8 + lr, g, q, af[i][0], af[i][1]);
            } else {
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int ip = 0; ip < MI / 2; ip++)
                        gemm_load_frag_mpair<BM>(at, wm * (8 * MI) + 16 * ip + 2 * lr, 8 * g + 2 * q + t, af[2 * ip][t],
                                                 af[2 * ip + 1][t]);
            }
            if (B_KMAJ) {
#pragma unroll
                for (int j = 0; j < NJ; j++)
                    gemm_load_frag<true, BN>(bt, wn * (8 * NJ) + j * 8 + lr, g, q, bf[j][0], bf[j][1]);
            } else {
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int jp = 0; jp < NJ / 2; jp++)
                        gemm_load_frag_mpair<BN>(bt, wn * (8 * NJ) + 16 * jp + 2 * lr, 8 * g + 2 * q + t, bf[2 * jp][t],
                                                 bf[2 * jp + 1][t]);
            }
#pragma unroll
            for (int t = 0; t < 2; t++)
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i][t], bf[j][t]);
        }
    }
    cp_async_wait<0>();

    // epilogue: every lane owns pairs of adjacent entries C[row][col], C[row][col + 1] (col even)
    const bool lower = p.flags & GEMM_LOWER;
    const bool beta0 = p.flags & GEMM_BETA0;
    const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    auto store_pair = [&](double *crow, int row, int col, double a0, double a1) {
        const int cmax = lower ? min(p.N, row + 1) : p.N;  // exclusive bound on valid columns
        if (col >= cmax) return;
        double v0 = p.alpha * a0, v1 = p.alpha * a1;
        if (p.scale) {
            const double sr = p.scale[row];
            v0 = (v0 * p.scale[col]) * sr;
            if (col + 1 < cmax) v1 = (v1 * p.scale[col + 1]) * sr;
        }
        if (vec_ok && col + 1 < cmax) {
            double2 *cp = reinterpret_cast<double2 *>(crow + col);
            if (!beta0) {
                double2 o = *cp;
                v0 += o.x;
                v1 += o.y;
            }
            *cp = make_double2(v0, v1);
        } else {
            if (!beta0) v0 += crow[col];
            crow[col] = v0;
            if (col + 1 < cmax) {
                if (!beta0) v1 += crow[col + 1];
                crow[col + 1] = v1;
            }
        }
        if (p.mir.n) {
            const int64_t off = (int64_t)row * p.mir.ld + col;
            if (col + 1 < cmax && !(p.mir.ld & 1)) {  // (destinations are 16-byte aligned: checked by the launcher)
                gemm_mirror_store2(p.mir, off, v0, v1);
            } else {
                gemm_mirror_store(p.mir, off, v0);
                if (col + 1 < cmax) gemm_mirror_store(p.mir, off + 1, v1);
            }
        }
    };
#pragma unroll
    for (int i = 0; i < MI; i++) {
        // row held by lane-row lr of MMA tile i (see gemm_load_frag_mpair for the m-major map)
        const int row = m0 + wm * (8 * MI) + (A_KMAJ ? i * 8 + lr : 16 * (i >> 1) + 2 * lr + (i & 1));
        if (row >= p.M) continue;
        double *crow = p.C + (int64_t)row * p.ldc;
        if (B_KMAJ) {
#pragma unroll
            for (int j = 0; j < NJ; j++)
                store_pair(crow, row, n0 + wn * (8 * NJ) + j * 8 + 2 * q, acc[i][j][0], acc[i][j][1]);
        } else {
            // tile-local column c of MMA tile 2p (+1) is matrix column 16p + 2c (+1): the lane's columns 2q, 2q + 1 of the
            // two tiles interleave into four adjacent matrix columns 16p + 4q .. + 3
#pragma unroll
            for (int jp = 0; jp < NJ / 2; jp++) {
                const int col = n0 + wn * (8 * NJ) + 16 * jp + 4 * q;
                store_pair(crow, row, col, acc[i][2 * jp][0], acc[i][2 * jp + 1][0]);
                store_pair(crow, row, col + 2, acc[i][2 * jp][1], acc[i][2 * jp + 1][1]);
            }
        }
    }
}

}  // namespace lgp
