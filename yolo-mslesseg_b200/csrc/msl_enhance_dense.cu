// Dense HE + CLAHE + GC + LT over PNG-oriented uint8 slice stacks (the per-slice stage of msl_enhance_volumes).
//
// Input : the normalised slices U that norm_scatter staged, PNG orientation P[r][c] (c fastest),
//         slice pitch a multiple of 16 bytes.  P[r, c] = G[c, cols-1-r]: P's column index c is the
//         slice ROW a (CLAHE's y, tile height th) and P's row index r mirrors the slice COLUMN
//         b = cols-1-r (CLAHE's x, tile width tw).  The kernel works in P coordinates end to end,
//         so neither the load nor the store transposes anything.
// Output: any of HE (E3, reference utils/mejora_imagen.py:52-67 == cv2.equalizeHist), CLAHE (E4, :91-117 ==
//         LUT_OUT[clahe(LUT_L[u])], OpenCV imgproc/clahe.cpp), GC (E5, :139-151) and LT (E6, :166-184) slices,
//         densely packed, same orientation.  One CTA per slice; all enhancements share one shared-memory copy.
//
// The kernel moves 2-5 B per pixel but is bound by shared-memory latency and instruction issue, so:
//  * Tile LUTs are built by autonomous warps (round 2): a warp pulls a tile from a queue, counts the tile's pixels into its
//    OWN 1 KB histogram (32-bit bins over u, background never counted), adds it to HE's histogram, adds the
//    BORDER_REFLECT_101 padding, folds the u-bins into L-bins (LUT_L is monotone), clips, redistributes and writes the
//    256-byte LUT - no block barrier between those steps, tiles without brain take a per-plane constant LUT, and the
//    histogram scratch is 16 KB instead of 64 KB.  Round 1 ran these as five block-wide phases (a sixth of the kernel was
//    barrier waits).
//  * Everything that depends only on the plane geometry (fold table, blend weights / offsets, the blank-tile LUT, the
//    blank-slice constant) is computed once per launch by dense_tables_kernel and copied into shared memory.
//  * HE / GC / LT are applied through one packed 32-bit table (one lookup per pixel for all three) and transposed
//    into three output words with PRMT.
//  * CLAHE tile LUTs are composed with LUT_L once per slice into "pair tables": for each tile row and gray the 9
//    horizontally adjacent (left, right) LUT byte pairs, 20 B per gray, so a pixel's four corners are two 16-bit
//    loads.  Blend weights and pair offsets per P row / column are tabulated, bytes are widened to float with
//    PRMT + magic subtract, and round-half-even goes through the 1.5*2^23 magic add instead of F2I.
//  * A slice that normalised to all zeros (outside the brain) skips every phase: outputs are table[0] constants.
#include "msl_common.cuh"
#include "msl_kernels.h"
#include <cstring>

namespace msl {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

struct DenseParams {
    const uint8_t* U;
    size_t u_pitch;            // bytes between input slices (multiple of 16)
    uint8_t* out_he;           // each may be NULL
    uint8_t* out_clahe;
    uint8_t* out_gc;
    uint8_t* out_lt;
    size_t out_pitch;          // bytes between output slices (= npx)
    const uint8_t* tables;
    const uint8_t* ptabs;      // per-plane tables written by dense_tables_kernel (CLAHE only)
    int rows, cols;            // slice orientation (G); P is cols x rows
    int th, tw, clip;
    float lut_scale;
};

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
    return p;
}

// shared memory map (bytes)
constexpr int kOffLutL = 0;        // [256] LUT_L
constexpr int kOffLutOut = 256;    // [256] LUT_OUT
constexpr int kOffT3 = 512;        // u32[256] he | gc << 8 | lt << 16
constexpr int kOffHeHist = 1536;   // u32[256]
constexpr int kOffMisc = 2560;     // int[32]
constexpr int kOffPT = 2688;       // per-plane tables (copy of the global block, layout below)
// per-plane table block (dense_tables_kernel); offsets relative to the block
constexpr int kPTFold = 0;         // u32[256] byte offsets of the (up to) two u-bins that fold into L-bin L: lo16 | hi16
constexpr int kPTUstart = 1024;    // u16[257] first u whose LUT_L[u] >= L   (514 B -> padded to 528)
constexpr int kPTProto = 1552;     // u8[256] tile LUT of a tile without brain: histogram {LUT_L[0]: th * tw}
constexpr int kPTMisc = 1808;      // int[4]: [0] long folds (some L-bin sums more than two u-bins), [1] CLAHE value of a blank slice
constexpr int kPTW = 1824;         // xw[cols] f32 | xo[cols] u32 | yw[rows] f32 | yo[rows] u32
constexpr int kPairTy = 256 * 20;          // bytes of pair tables per tile row: 256 grays x (9 pairs x 2 B, padded to 20)
constexpr int kPairStride = kPairTy + 1024 + 4;   // each tile row's pair tables are followed by its background row TZ[ty][r] (<= 256 floats);
                                                  // + 4: consecutive tile rows start one bank apart (a warp reads TZ of two tile rows at one r)
constexpr int kPairBytes = (8 * kPairStride + 15) & ~15;   // 49,184
constexpr int kRBytes = kPairBytes + 64 * 256;  // R: pair tables + background rows, then the 64 tile LUTs Tc[tile][L]; while the tile LUTs
                                                // are built the head of R holds one 1 KB histogram per warp
static_assert(kWarps * 1024 <= kPairBytes, "histogram scratch must fit in front of the tile LUTs");

// tiles in centre-first order: the tiles with the most brain are pulled first, the cheap border tiles fill the tail
__constant__ uint8_t kTileOrder[64] = {27, 28, 35, 36, 19, 20, 26, 29, 34, 37, 43, 44, 18, 21, 42, 45, 11, 12, 25, 30, 33, 38,
                                       51, 52, 10, 13, 17, 22, 41, 46, 50, 53, 3,  4,  9,  14, 24, 31, 32, 39, 49, 54, 59, 60,
                                       2,  5,  16, 23, 40, 47, 58, 61, 1,  6,  8,  15, 48, 55, 57, 62, 0,  7,  56, 63};

// floor(a / b) for 0 <= a < 2^16, 1 <= b <= 256 without the integer-division sequence: the approximate quotient is biased
// up by 4e-6 (more than its error, less than the 1 / b gap below the next integer).
__device__ __forceinline__ int small_div(int a, int b) {
    return (int)__fmul_rn(__fdividef((float)a, (float)b), 1.000004f);
}

// Packed FP32 pairs (Blackwell FMUL2 / FADD2): two IEEE round-to-nearest operations per issue slot.  ptxas contracts a
// mul.f32x2 feeding an add.f32x2 into FFMA2 even with -fmad=false, which would drop a rounding step, so add2 is only used
// on operands that are not products (the exact magic-subtract conversions); sums of products stay scalar.
__device__ __forceinline__ void mul2(float& o0, float& o1, float a0, float a1, float b0, float b1) {
    unsigned long long a, b, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(r));
}
__device__ __forceinline__ void add2(float& o0, float& o1, float a0, float a1, float b0, float b1) {
    unsigned long long a, b, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(r));
}

// Up to three stacks (the three planes of a chunk of volumes) in one launch: one grid instead of three means one tail
// instead of three.  CTA b works on slice b - first[k] of stack k.
struct DenseLaunch {
    DenseParams plane[3];
    int first[4];
};

// OpenCV CLAHE_CalcLut_Body for one tile whose 256 L-bins are spread over a warp, 8 consecutive bins per lane: clip,
// redistribute the excess (residualStep walk), CDF, LUT = saturate_cast<uchar>(cdf * lutScale).  Returns the lane's 8 LUT bytes.
__device__ __forceinline__ uint2 clip_cdf_tile(int (&hb)[8], int lane, int clip, float lut_scale) {
    int clipped = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (hb[k] > clip) { clipped += hb[k] - clip; hb[k] = clip; }
    clipped = warp_sum(clipped);
    const int rb = clipped >> 8;
    const int res = clipped & 255;
    // residual: bins 0, step, 2*step, ... (res of them) get one more; walk this lane's 8 bins without dividing per bin
    const int step = res > 0 ? small_div(256, res) : 256;
    const int base = lane * 8;
    int kn = small_div(base + step - 1, step);          // index of the first multiple of step that is >= base
    int nxt = kn * step;
    int run = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        hb[k] += rb;
        const bool hit = (nxt == base + k) && (kn < res);
        if (hit) { hb[k] += 1; nxt += step; ++kn; }
        run += hb[k];
        hb[k] = run;
    }
    const int excl = warp_incl_scan(run, lane) - run;
    // saturate_cast<uchar>(cdf * lutScale): the product lies in [0, 255.0001], so int -> float and round-half-even both
    // go through magic adds and the low byte of the sum is the result
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float cdf = __fsub_rn(__uint_as_float(0x4b000000u + (uint32_t)(hb[k] + excl)), 8388608.0f);
        o[k] = __float_as_uint(__fadd_rn(__fmul_rn(cdf, lut_scale), 12582912.0f));
    }
    const uint32_t lo = __byte_perm(__byte_perm(o[0], o[1], 0x0040), __byte_perm(o[2], o[3], 0x0040), 0x5410);
    const uint32_t hi = __byte_perm(__byte_perm(o[4], o[5], 0x0040), __byte_perm(o[6], o[7], 0x0040), 0x5410);
    return make_uint2(lo, hi);
}

// Everything that depends only on the plane (geometry + LUT_L): one CTA of 288 threads per stack, once per launch.
__global__ void __launch_bounds__(288) dense_tables_kernel(const __grid_constant__ DenseLaunch L) {
    __shared__ uint8_t lutl[256];
    __shared__ uint16_t ustart[257];
    const DenseParams& p = L.plane[blockIdx.x];
    uint8_t* tb = const_cast<uint8_t*>(p.ptabs);
    const int tid = threadIdx.x, lane = tid & 31;
    const int rows = p.rows, cols = p.cols, th = p.th, tw = p.tw;
    if (tid < 256) lutl[tid] = __ldg(p.tables + MSL_TAB_LUT_L + tid);
    if (tid == 0) reinterpret_cast<int*>(tb + kPTMisc)[0] = 0;
    __syncthreads();
    if (tid <= 256) {
        // LUT_L is monotone: the u-bins that fold into L-bin L are [ustart[L], ustart[L+1])
        int lo = 0, hi = 256;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)lutl[mid] >= tid) hi = mid; else lo = mid + 1; }
        ustart[tid] = (uint16_t)lo;
        reinterpret_cast<uint16_t*>(tb + kPTUstart)[tid] = (uint16_t)lo;
    }
    __syncthreads();
    if (tid < 256) {
        // fold table: L-bin tid sums u-bins [ustart, ustart + n).  n <= 2 for cv2's LUT_L: keep two byte offsets, the
        // unused ones pointing at u-bin 0, which is never counted and stays zero.  n > 2 (other tables): generic loop.
        const int u0 = ustart[tid], n = (int)ustart[tid + 1] - u0;
        reinterpret_cast<uint32_t*>(tb + kPTFold)[tid] = (uint32_t)(n > 0 ? u0 * 4 : 0) | ((uint32_t)(n > 1 ? (u0 + 1) * 4 : 0) << 16);
        if (n > 2) reinterpret_cast<int*>(tb + kPTMisc)[0] = 1;
    }
    // interpolation tables (OpenCV CLAHE_Interpolation_Body): blend weight + table offsets per P row / P column.
    //   P row r    (slice column b): weight xa, pair slot j = floor(txf) + 1 in [0, 8]  -> byte offset 2*j
    //   P column c (slice row a)   : weight ya, tile rows ty1 / ty2
    {
        float* xw = reinterpret_cast<float*>(tb + kPTW);
        uint32_t* xo = reinterpret_cast<uint32_t*>(xw + cols);
        float* yw = reinterpret_cast<float*>(xo + cols);
        uint32_t* yo = reinterpret_cast<uint32_t*>(yw + rows);
        const float inv_tw = __fdiv_rn(1.0f, (float)tw), inv_th = __fdiv_rn(1.0f, (float)th);
        for (int r = tid; r < cols; r += blockDim.x) {
            const int b = cols - 1 - r;
            const float txf = __fsub_rn(__fmul_rn((float)b, inv_tw), 0.5f);
            const int t1 = (int)floorf(txf);
            xw[r] = __fsub_rn(txf, (float)t1);
            xo[r] = (uint32_t)(2 * (min(max(t1, -1), 7) + 1));
        }
        for (int a = tid; a < rows; a += blockDim.x) {
            const float tyf = __fsub_rn(__fmul_rn((float)a, inv_th), 0.5f);
            const int t1 = (int)floorf(tyf), t2 = t1 + 1;
            yw[a] = __fsub_rn(tyf, (float)t1);
            yo[a] = (uint32_t)max(t1, 0) | ((uint32_t)min(t2, 7) << 16);
        }
    }
    // A tile without brain holds th * tw pixels of gray 0: its LUT is a constant of the plane.  A blank slice is 64 such
    // tiles; the blend of four equal tile values z is z after rounding, so every pixel is LUT_OUT[T[LUT_L[0]]].
    if (tid < 32) {
        const int L0 = lutl[0];
        int hb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hb[k] = (lane * 8 + k == L0) ? th * tw : 0;
        const uint2 t = clip_cdf_tile(hb, lane, p.clip, p.lut_scale);
        reinterpret_cast<uint2*>(tb + kPTProto)[lane] = t;
        if (lane == (L0 >> 3)) {
            const uint32_t w = (L0 & 4) ? t.y : t.x;
            reinterpret_cast<int*>(tb + kPTMisc)[1] = (int)__ldg(p.tables + MSL_TAB_LUT_OUT + ((w >> (8 * (L0 & 3))) & 0xffu));
        }
    }
}

// W_CT: the P row length (= slice rows) as a compile-time constant for the MSLesSeg planes (182, 218), 0 = run time.  With
// it the row-strided accesses of the histogram and blend loops (px[i * W], op[i * W]) are immediate offsets.
template <bool DO_CLAHE, int W_CT>
__device__ __forceinline__ void dense_slice(const DenseParams& p, const size_t s, uint8_t* smem) {
    uint8_t* lutl = smem + kOffLutL;
    uint8_t* lutout = smem + kOffLutOut;
    uint32_t* t3 = reinterpret_cast<uint32_t*>(smem + kOffT3);
    unsigned* he_hist = reinterpret_cast<unsigned*>(smem + kOffHeHist);
    int* misc = reinterpret_cast<int*>(smem + kOffMisc);       // [0] i0, [1..8] warp scan totals, [22] tile queue
    uint8_t* pt = smem + kOffPT;
    const uint32_t* fold = reinterpret_cast<const uint32_t*>(pt + kPTFold);
    const uint16_t* ustart = reinterpret_cast<const uint16_t*>(pt + kPTUstart);
    const int* ptmisc = reinterpret_cast<const int*>(pt + kPTMisc);
    const int rows = p.rows, cols = p.cols, npx = rows * cols;
    const int W = W_CT ? W_CT : rows;                            // P row length
    const float* xw = reinterpret_cast<const float*>(pt + kPTW);       // indexed by P row r   (slice column b = cols-1-r)
    const uint32_t* xo = reinterpret_cast<const uint32_t*>(xw + cols);
    const float* yw = reinterpret_cast<const float*>(xo + cols);         // indexed by P column c (slice row a = c)
    const uint32_t* yo = reinterpret_cast<const uint32_t*>(yw + rows);
    // (offsets, not pointer casts: an integer round trip would make the compiler fall back to generic LD/ST)
    const int ptab_bytes = DO_CLAHE ? ((kPTW + (rows + cols) * 8 + 15) & ~15) : 0;
    const int offR = kOffPT + ptab_bytes;
    uint8_t* R = smem + offR;
    uint8_t* su = R + (DO_CLAHE ? kRBytes : 0);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* in = p.U + s * p.u_pitch;
    const bool want_he = p.out_he != nullptr, want_lut = want_he || p.out_gc || p.out_lt;

    // ---------------------------------------------------------------- load
    // The slice's 128-bit loads are issued first; the table copies overlap their latency.
    constexpr int kMaxVec = 5;                                   // 5 x 16 B per thread in flight (covers 40 KB slices)
    const int nvec = npx >> 4;
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    uint4 pre[kMaxVec];
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
        const int q = tid + j * kThreads;
        pre[j] = q < nvec ? __ldg(in4 + q) : make_uint4(0, 0, 0, 0);
    }
    if (tid < 128) reinterpret_cast<uint32_t*>(lutl)[tid] = __ldg(reinterpret_cast<const uint32_t*>(p.tables) + tid);  // LUT_L + LUT_OUT
    if (tid < 256) he_hist[tid] = 0;
    if (tid == 0) { misc[0] = 256; misc[22] = 0; }
    if (DO_CLAHE) {
        // the plane's table block goes global -> shared without passing through registers (cp.async, SASS LDGSTS)
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(pt);
        for (int q = tid; q < (ptab_bytes >> 4); q += kThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * (uint32_t)q), "l"(p.ptabs + 16 * (size_t)q) : "memory");
    }
    uint32_t anynz = 0;
    {
        uint4* su4 = reinterpret_cast<uint4*>(su);
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int q = tid + j * kThreads;
            if (q < nvec) { su4[q] = pre[j]; anynz |= pre[j].x | pre[j].y | pre[j].z | pre[j].w; }
        }
        for (int q = tid + kMaxVec * kThreads; q < nvec; q += kThreads) { const uint4 v = __ldg(in4 + q); su4[q] = v; anynz |= v.x | v.y | v.z | v.w; }
        for (int o = (nvec << 4) + tid; o < npx; o += kThreads) { const uint32_t b = __ldg(in + o); su[o] = (uint8_t)b; anynz |= b; }
    }
    if (DO_CLAHE) asm volatile("cp.async.wait_all;" ::: "memory");
    // Blank slice (the skull-stripped volumes have ~15 % of them per plane): every output is a constant.
    // HE: one populated bin -> that bin's index (0); GC_T[0]; LT_T[.][0]; CLAHE: the per-plane constant of dense_tables_kernel.
    if (!__syncthreads_or(anynz != 0)) {
        const uint32_t c_he = 0, c_gc = __ldg(p.tables + MSL_TAB_GC), c_lt = __ldg(p.tables + MSL_TAB_LT + 255 * 256);
        const uint32_t c_cl = DO_CLAHE ? (uint32_t)ptmisc[1] : 0u;
        uint8_t* outs[4] = {p.out_he, p.out_clahe, p.out_gc, p.out_lt};
        const uint32_t cv[4] = {c_he, c_cl, c_gc, c_lt};
        const int nw0 = npx >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!outs[j]) continue;
            uint8_t* out = outs[j] + s * p.out_pitch;
            uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
            const uint32_t w4 = cv[j] * 0x01010101u;
            for (int q = tid; q < nw0; q += kThreads) out32[q] = w4;
            for (int o = (nw0 << 2) + tid; o < npx; o += kThreads) out[o] = (uint8_t)cv[j];
        }
        return;
    }

    const int th = p.th, tw = p.tw;
    uint8_t* Tc = R + kPairBytes;                             // [64][256] tile LUTs
    const uint32_t* su32 = reinterpret_cast<const uint32_t*>(su);
    const int nw = npx >> 2;

    if (DO_CLAHE) {
        // ------------------------------------------------------------ tile LUTs, one warp per tile, no block barrier inside
        // (OpenCV CLAHE_CalcLut_Body).  A lane owns one P column of the tile (slice row a -> up to th
        // lanes busy) and walks the tile's P rows eight at a time.  Background pixels (u == 0) are never counted: every
        // padded tile holds th * tw pixels, so bin 0 is recovered by subtraction (and HE's bin 0 in its CDF).
        unsigned* ht = reinterpret_cast<unsigned*>(R + warp * 1024);
        const uint8_t* hub = reinterpret_cast<const uint8_t*>(ht);
        const bool long_fold = ptmisc[0] != 0;
        const int L0 = lutl[0], area = th * tw, clip = p.clip;
        for (;;) {
            int q = 0;
            if (lane == 0) q = atomicAdd(&misc[22], 1);
            q = __shfl_sync(FULL, q, 0);
            if (q >= 64) break;
            const int tile = kTileOrder[q], ty = tile >> 3, tx = tile & 7;
            reinterpret_cast<uint4*>(ht)[lane] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(ht)[lane + 32] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            const int a_lo = ty * th, a_hi = a_lo + th;                      // slice rows of the tile (padded extent)
            const int b_lo = tx * tw, b_hi = b_lo + tw;                      // slice columns of the tile (padded extent)
            const int r0 = max(0, cols - b_hi), r_end = cols - b_lo;         // real P rows of the tile (empty if r0 >= r_end)
            uint32_t seen = 0;
            // real pixels
            for (int a0 = a_lo; a0 < min(a_hi, rows); a0 += 32) {
                const int c = a0 + lane;
                if (c >= min(a_hi, rows)) continue;
                const uint8_t* px = su + r0 * W + c;
                int r = r0;
                for (; r + 7 < r_end; r += 8, px += 8 * W) {       // eight independent loads in flight
                    uint32_t v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = px[i * W];
                    const uint32_t any = v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7];
                    if (any == 0) continue;
                    seen |= any;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (v[i]) atomicAdd(&ht[v[i]], 1u);
                }
                for (; r + 3 < r_end; r += 4, px += 4 * W) {
                    const uint32_t v0 = px[0], v1 = px[W], v2 = px[2 * W], v3 = px[3 * W];    // loads first, atomics after
                    if ((v0 | v1 | v2 | v3) == 0) continue;
                    seen |= v0 | v1 | v2 | v3;
                    if (v0) atomicAdd(&ht[v0], 1u);
                    if (v1) atomicAdd(&ht[v1], 1u);
                    if (v2) atomicAdd(&ht[v2], 1u);
                    if (v3) atomicAdd(&ht[v3], 1u);
                }
                for (; r < r_end; ++r, px += W) {
                    const uint32_t v = px[0];
                    if (v) { atomicAdd(&ht[v], 1u); seen = 1; }
                }
            }
            bool nz = __any_sync(FULL, seen != 0);
            __syncwarp();                                         // the warp's shared-memory atomics are ordered before the reads below
            // HE's histogram = sum of the tile histograms of the REAL pixels (before the CLAHE padding is added)
            if (want_he && nz) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned h = ht[lane + 32 * k];
                    if (h) atomicAdd(&he_hist[lane + 32 * k], h);
                }
                __syncwarp();                                     // ... and those reads before the padding is added
            }
            // BORDER_REFLECT_101 padding (OpenCV pads bottom / right up to the 8x8 tile grid): padded slice rows a >= rows
            // over every column of the tile, and real rows over the padded columns b >= cols
            if (a_hi > rows || b_hi > cols) {
                uint32_t seen2 = 0;
                // padded slice rows (at most 8 of them): lanes run along the tile's slice columns b, padded ones included
                for (int ap = max(a_lo, rows); ap < a_hi; ++ap) {
                    const int asrc = reflect101(ap, rows);
                    for (int b0 = b_lo; b0 < b_hi; b0 += 32) {
                        const int b = b0 + lane;
                        if (b >= b_hi) continue;
                        const uint32_t v = su[(cols - 1 - (b < cols ? b : reflect101(b, cols))) * W + asrc];
                        if (v) { atomicAdd(&ht[v], 1u); seen2 = 1; }
                    }
                }
                // padded slice columns (at most 8) over the real rows: lanes run along the tile's slice rows a
                for (int bp = max(b_lo, cols); bp < b_hi; ++bp) {
                    const uint8_t* prow = su + (cols - 1 - reflect101(bp, cols)) * W;
                    for (int a0 = a_lo; a0 < min(a_hi, rows); a0 += 32) {
                        const int a = a0 + lane;
                        if (a >= min(a_hi, rows)) continue;
                        const uint32_t v = prow[a];
                        if (v) { atomicAdd(&ht[v], 1u); seen2 = 1; }
                    }
                }
                nz |= __any_sync(FULL, seen2 != 0);
            }
            __syncwarp();
            if (!nz) {                                            // no brain in this tile: the plane's constant LUT
                reinterpret_cast<uint2*>(Tc + tile * 256)[lane] = reinterpret_cast<const uint2*>(pt + kPTProto)[lane];
                continue;
            }
            // fold the u-bins into L-bins, 8 L-bins per lane
            int hb[8];
            if (!long_fold) {
                const uint4 f0 = reinterpret_cast<const uint4*>(fold)[lane * 2], f1 = reinterpret_cast<const uint4*>(fold)[lane * 2 + 1];
                const uint32_t f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    hb[k] = (int)(*reinterpret_cast<const unsigned*>(hub + (f[k] & 0xffffu)) + *reinterpret_cast<const unsigned*>(hub + (f[k] >> 16)));
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int Lb = lane * 8 + k;
                    int acc = 0;
                    for (int u = ustart[Lb]; u < (int)ustart[Lb + 1]; ++u) acc += (int)ht[u];
                    hb[k] = acc;
                }
            }
            {
                // the tile's background pixels were never counted: th * tw minus everything else, into L-bin LUT_L[0]
                int tot = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) tot += hb[k];
                const int zeros = area - warp_sum(tot);
                if (L0 == 0) { if (lane == 0) hb[0] += zeros; }
                else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (lane * 8 + k == L0) hb[k] += zeros;
                }
            }
            reinterpret_cast<uint2*>(Tc + tile * 256)[lane] = clip_cdf_tile(hb, lane, clip, p.lut_scale);
            __syncwarp();                                         // the scratch histogram is cleared for the next tile
        }
    } else if (want_he) {
        // HE without CLAHE: plain 256-bin histogram, zero words skipped
        int zeros = 0;
        for (int q = tid; q < nw; q += kThreads) {
            const uint32_t w = su32[q];
            if (w == 0) { zeros += 4; continue; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t b = (w >> (8 * k)) & 0xff;
                if (b) atomicAdd(&he_hist[b], 1u); else ++zeros;
            }
        }
        for (int o = (nw << 2) + tid; o < npx; o += kThreads) { if (su[o]) atomicAdd(&he_hist[su[o]], 1u); else ++zeros; }
        zeros = warp_sum(zeros);
        if (lane == 0 && zeros) atomicAdd(&he_hist[0], (unsigned)zeros);
    }
    __syncthreads();

    // ---------------------------------------------------------------- HE CDF -> LUT; packed HE | GC | LT table
    // Built by warps 0-7 (one gray level per thread).  With CLAHE on, the other warps do not wait for it: they start on
    // the pair tables below, and the table is published by the barrier behind those.
    auto build_t3 = [&](bool active, auto sync8) {           // active: tid < 256; sync8: a barrier over all callers
        // (with CLAHE on, he_hist[0] is still empty: the background count is npx minus everything else)
        int h = 0, c = 0;
        if (want_he && active) {
            h = (int)he_hist[tid];
            c = warp_incl_scan(h, lane);
            if (lane == 31) misc[1 + warp] = c;
            if (h > 0) atomicMin(&misc[0], tid);
        }
        sync8();
        if (!active) return;
        uint32_t he = 0;
        if (want_he) {
            int total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const int m = misc[1 + w]; total += m; if (w < warp) c += m; }
            const int zeros = npx - total;                       // uncounted background pixels (0 without CLAHE)
            c += zeros;
            const int i0 = zeros > 0 ? 0 : misc[0];
            const int h0 = zeros > 0 ? zeros + (int)he_hist[0] : (int)he_hist[i0];
            if (h0 == npx) he = (uint32_t)i0;
            else if (tid <= i0) he = 0;
            else he = sat_u8_rn(__fmul_rn((float)(c - h0), __fdiv_rn(255.0f, (float)(npx - h0))));
        }
        // LT: E1 maps the slice maximum to exactly 255 whenever ptp > 0, so the table row is 255; a blank slice is
        // all zeros and LT_T[255][0] == LT_T[0][0] == 0.
        const uint32_t gc = __ldg(p.tables + MSL_TAB_GC + tid), lt = __ldg(p.tables + MSL_TAB_LT + 255 * 256 + tid);
        t3[tid] = he | (gc << 8) | (lt << 16);
    };
    // one lookup per pixel for HE, GC and LT; the 4 x 3 bytes of a word are transposed into three output words
    auto map_t3 = [&]() {
        uint32_t* o_he = reinterpret_cast<uint32_t*>(p.out_he ? p.out_he + s * p.out_pitch : nullptr);
        uint32_t* o_gc = reinterpret_cast<uint32_t*>(p.out_gc ? p.out_gc + s * p.out_pitch : nullptr);
        uint32_t* o_lt = reinterpret_cast<uint32_t*>(p.out_lt ? p.out_lt + s * p.out_pitch : nullptr);
        const uint32_t z = t3[0];
        const uint32_t z_he = (z & 0xffu) * 0x01010101u, z_gc = ((z >> 8) & 0xffu) * 0x01010101u, z_lt = ((z >> 16) & 0xffu) * 0x01010101u;
        for (int q0 = warp * 32; q0 < nw; q0 += kThreads) {
            const int q = q0 + lane;
            const bool act = q < nw;
            const uint32_t w = act ? su32[q] : 0u;
            if (!__any_sync(FULL, w != 0)) {                     // 32 background words (two thirds of them): constants
                if (act) {
                    if (o_he) o_he[q] = z_he;
                    if (o_gc) o_gc[q] = z_gc;
                    if (o_lt) o_lt[q] = z_lt;
                }
                continue;
            }
            if (!act) continue;
            uint32_t a0 = z, a1 = z, a2 = z, a3 = z;
            if (w) { a0 = t3[w & 0xff]; a1 = t3[(w >> 8) & 0xff]; a2 = t3[(w >> 16) & 0xff]; a3 = t3[w >> 24]; }
            // transpose the 4 x 3 bytes: byte j of every a_k -> output word j
            const uint32_t lo01 = __byte_perm(a0, a1, 0x5140), lo23 = __byte_perm(a2, a3, 0x5140);   // he0 he1 gc0 gc1 | he2 he3 gc2 gc3
            if (o_he) o_he[q] = __byte_perm(lo01, lo23, 0x5410);
            if (o_gc) o_gc[q] = __byte_perm(lo01, lo23, 0x7632);
            if (o_lt) o_lt[q] = __byte_perm(__byte_perm(a0, a1, 0x0062), __byte_perm(a2, a3, 0x0062), 0x5410);
        }
        for (int o = (nw << 2) + tid; o < npx; o += kThreads) {
            const uint32_t a0 = t3[su[o]];
            if (p.out_he) p.out_he[s * p.out_pitch + o] = (uint8_t)a0;
            if (p.out_gc) p.out_gc[s * p.out_pitch + o] = (uint8_t)(a0 >> 8);
            if (p.out_lt) p.out_lt[s * p.out_pitch + o] = (uint8_t)(a0 >> 16);
        }
    };
    if (!DO_CLAHE) {
        if (want_lut) {
            build_t3(tid < 256, [] { __syncthreads(); });
            __syncthreads();
            map_t3();
        }
        return;
    }
    if (want_lut && warp < 8) build_t3(true, [] { asm volatile("bar.sync 1, 256;" ::: "memory"); });

    const bool use_tz = cols <= (kPairStride - kPairTy) / 4;
    // Pair tables: PT[ty][u][j] = (T[ty][tx1][LUT_L[u]], T[ty][tx2][LUT_L[u]]) as one 16-bit entry for the nine
    // horizontal neighbour pairs (tx1, tx2) = (0,0), (0,1), ..., (6,7), (7,7).  One 16-bit read fetches both operands
    // of a horizontal blend; a gray level's nine entries take 20 bytes (5 words: odd stride -> spread over the banks).
    // Thread = one gray level u and four tile rows: 8 byte reads (neighbouring u -> neighbouring L: conflict-free)
    // and five 32-bit stores per tile row.  (The histogram scratch at the head of R is dead since the barrier above.)
    {
        const int uv = tid & 255, Lv = lutl[uv];
        for (int ty = tid >> 8; ty < 8; ty += kThreads / 256) {
            uint32_t t[8];
#pragma unroll
            for (int tx = 0; tx < 8; ++tx) t[tx] = Tc[(ty * 8 + tx) * 256 + Lv];
            uint32_t* dst = reinterpret_cast<uint32_t*>(R + ty * kPairStride + uv * 20);
            // pairs j = 0..8: (t0,t0) (t0,t1) (t1,t2) ... (t6,t7) (t7,t7), two 16-bit pairs per word
            dst[0] = (t[0] | (t[0] << 8)) | ((t[0] | (t[1] << 8)) << 16);
            dst[1] = (t[1] | (t[2] << 8)) | ((t[2] | (t[3] << 8)) << 16);
            dst[2] = (t[3] | (t[4] << 8)) | ((t[4] | (t[5] << 8)) << 16);
            dst[3] = (t[5] | (t[6] << 8)) | ((t[6] | (t[7] << 8)) << 16);
            dst[4] = (t[7] | (t[7] << 8));
        }
        // Background rows: for u == 0 the horizontal half of the blend depends only on (tile row, P row).  TZ[ty][r] holds it
        // (the same two products and sum the per-pixel path computes), so a background pixel needs two loads and the
        // vertical half.  Lives behind each tile row's pair tables (slices up to 256 P rows; others take the pixel path).
        if (use_tz) {
            const int L0 = lutl[0];
            for (int r = tid; r < cols; r += kThreads) {
                const int j = (int)(xo[r] >> 1), tx1 = max(j - 1, 0), tx2 = min(j, 7);
                const float xa = xw[r], xa1 = __fsub_rn(1.0f, xa);
#pragma unroll
                for (int ty = 0; ty < 8; ++ty) {
                    const float lo = (float)Tc[(ty * 8 + tx1) * 256 + L0], hi = (float)Tc[(ty * 8 + tx2) * 256 + L0];
                    reinterpret_cast<float*>(R + ty * kPairStride + kPairTy)[r] = __fadd_rn(__fmul_rn(lo, xa1), __fmul_rn(hi, xa));
                }
            }
        }
    }
    __syncthreads();
    if (want_lut) map_t3();                                  // HE | GC | LT outputs (the table was published by the barrier above)

    // ---------------------------------------------------------------- CLAHE: bilinear blend + LUT_OUT
    // A warp owns 32 consecutive P columns (slice rows) over a band of P rows: the vertical weight / offsets stay in
    // registers, the horizontal ones are warp-uniform reads, pixel reads and writes are conflict-free.
    {
        uint8_t* out = p.out_clahe + s * p.out_pitch;     // a warp's 32 results are one 32-byte row segment: stored straight to global
        const int nchunk = (W + 31) >> 5;
        const int nband = 8, band_rows = (cols + nband - 1) / nband;
        for (int task = warp; task < nchunk * nband; task += kWarps) {
            const int cc = task % nchunk, band = task / nchunk;
            const int c = cc * 32 + lane;
            const unsigned amask = __ballot_sync(0xffffffffu, c < W);
            if (c >= W) continue;
            const float ya = yw[c], ya1 = __fsub_rn(1.0f, ya);
            const uint32_t yoff = yo[c];
            const uint8_t* P1 = R + (yoff & 0xffff) * kPairStride;
            const uint8_t* P2 = R + (yoff >> 16) * kPairStride;
            const float* tz1 = reinterpret_cast<const float*>(P1 + kPairTy);
            const float* tz2 = reinterpret_cast<const float*>(P2 + kPairTy);
            const int r_end = min(cols, (band + 1) * band_rows);
            auto blend = [&](int r, uint32_t v) -> uint8_t {
                const float xa = xw[r], xa1 = __fsub_rn(1.0f, xa);
                const uint32_t off = 20u * v + xo[r];
                const uint32_t h1 = *reinterpret_cast<const uint16_t*>(P1 + off);
                const uint32_t h2 = *reinterpret_cast<const uint16_t*>(P2 + off);
                // uint8 -> float without the conversion pipe: bits(2^23 + b) - 2^23 (PRMT builds the bits)
                float l11, l12, l21, l22;
                add2(l11, l21, __uint_as_float(__byte_perm(h1, 0x4b000000u, 0x7540)), __uint_as_float(__byte_perm(h2, 0x4b000000u, 0x7540)), -8388608.0f, -8388608.0f);
                add2(l12, l22, __uint_as_float(__byte_perm(h1, 0x4b000000u, 0x7541)), __uint_as_float(__byte_perm(h2, 0x4b000000u, 0x7541)), -8388608.0f, -8388608.0f);
                float a1, a2, b1, b2, t1, t2;
                mul2(a1, a2, l11, l21, xa1, xa1);
                mul2(b1, b2, l12, l22, xa, xa);
                const float top = __fadd_rn(a1, b1), bot = __fadd_rn(a2, b2);
                mul2(t1, t2, top, bot, ya1, ya);
                const float res = __fadd_rn(t1, t2);
                // cvRound (half-even) of a value in [0, 255.0001]: the low mantissa bits of res + 1.5 * 2^23
                return lutout[__float_as_uint(__fadd_rn(res, 12582912.0f)) & 0xffu];
            };
            // two rows per iteration: both pixels are read before either result is written back, so the two
            // dependent chains (pixel -> pair table -> blend -> LUT_OUT) overlap
            int r = band * band_rows;
            auto blend_bg = [&](int r) -> uint8_t {
                float t1, t2;
                mul2(t1, t2, tz1[r], tz2[r], ya1, ya);
                const float res = __fadd_rn(t1, t2);
                return lutout[__float_as_uint(__fadd_rn(res, 12582912.0f)) & 0xffu];
            };
            uint8_t* op = out + (unsigned)(r * W + c);        // running output pointer
            const unsigned Wu = (unsigned)W;
            const uint8_t* ip = su + r * W + c;
            // four rows per iteration: one vote decides whether all 32 x 4 pixels are background (the short path)
            for (; r + 3 < r_end; r += 4, op += 4 * Wu, ip += 4 * W) {
                const uint32_t v0 = ip[0], v1 = ip[W], v2 = ip[2 * W], v3 = ip[3 * W];
                uint8_t g0, g1, g2, g3;
                if (use_tz && !__any_sync(amask, (v0 | v1 | v2 | v3) != 0)) {
                    g0 = blend_bg(r); g1 = blend_bg(r + 1); g2 = blend_bg(r + 2); g3 = blend_bg(r + 3);
                } else {
                    g0 = blend(r, v0); g1 = blend(r + 1, v1);
                    g2 = blend(r + 2, v2); g3 = blend(r + 3, v3);
                }
                op[0] = g0; op[Wu] = g1; op[2 * Wu] = g2; op[3 * Wu] = g3;
            }
            for (; r + 1 < r_end; r += 2, op += 2 * Wu, ip += 2 * W) {
                const uint32_t v0 = ip[0], v1 = ip[W];
                uint8_t g0, g1;
                if (use_tz && !__any_sync(amask, (v0 | v1) != 0)) { g0 = blend_bg(r); g1 = blend_bg(r + 1); }   // warp-uniform
                else { g0 = blend(r, v0); g1 = blend(r + 1, v1); }
                op[0] = g0;
                op[Wu] = g1;
            }
            if (r < r_end) op[0] = blend(r, su[r * W + c]);
        }
    }
}

template <bool DO_CLAHE>
__global__ void __launch_bounds__(kThreads, 2) enhance_dense_kernel(const __grid_constant__ DenseLaunch L) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int plane_k = (int)blockIdx.x >= L.first[2] ? 2 : ((int)blockIdx.x >= L.first[1] ? 1 : 0);
    const DenseParams& p = L.plane[plane_k];
    const size_t s = (size_t)((int)blockIdx.x - L.first[plane_k]);
    if (p.rows == 182) dense_slice<DO_CLAHE, 182>(p, s, smem);
    else if (p.rows == 218) dense_slice<DO_CLAHE, 218>(p, s, smem);
    else dense_slice<DO_CLAHE, 0>(p, s, smem);
}

}  // namespace

size_t dense_u_pitch(int npx) { return ((size_t)npx + 15) & ~(size_t)15; }

// bytes of one plane's table block (dense_tables_kernel), a multiple of 16
size_t dense_tabs_bytes(int rows, int cols) { return ((size_t)kPTW + (size_t)(rows + cols) * 8 + 15) & ~(size_t)15; }

size_t dense_smem_bytes(int rows, int cols, bool clahe) {
    return (size_t)kOffPT + (clahe ? dense_tabs_bytes(rows, cols) + kRBytes : 0) + dense_u_pitch(rows * cols);
}

bool dense_supported(int rows, int cols, bool clahe) {
    const long long npx = (long long)rows * cols;
    return (npx % 4 == 0) && rows >= 2 && npx * rows < 0x100000000ll && rows + cols + 8 <= 2048 &&
           dense_smem_bytes(rows, cols, clahe) <= 227 * 1024;
}

int launch_enhance_dense_multi(const DensePlane* planes, int nplanes, const uint8_t* tables, void* tabs_ws, size_t tabs_ws_bytes,
                               cudaStream_t stream) {
    DenseLaunch L;
    memset(&L, 0, sizeof(L));
    int n = 0, total = 0;
    bool cl = false;
    size_t smem = 0, tabs_used = 0;
    for (int i = 0; i < nplanes; ++i) {
        const DensePlane& q = planes[i];
        if (q.nslices <= 0 || (!q.out_he && !q.out_clahe && !q.out_gc && !q.out_lt)) continue;
        if (n == 3) { set_error("enhance_dense: at most three stacks per launch"); return MSL_ERR_ARG; }
        const bool qcl = q.out_clahe != nullptr;
        if (n > 0 && qcl != cl) { set_error("enhance_dense: the stacks of one launch must agree on CLAHE"); return MSL_ERR_ARG; }
        cl = qcl;
        const size_t sm = dense_smem_bytes(q.rows, q.cols, cl);
        if (!dense_supported(q.rows, q.cols, cl) || (q.u_pitch & 15) || (reinterpret_cast<uintptr_t>(q.U) & 15) ||
            ((reinterpret_cast<uintptr_t>(q.out_he) | reinterpret_cast<uintptr_t>(q.out_clahe) | reinterpret_cast<uintptr_t>(q.out_gc) |
              reinterpret_cast<uintptr_t>(q.out_lt)) & 3)) {
            set_error("enhance_dense: unsupported geometry / alignment (%d x %d, %zu B smem)", q.rows, q.cols, sm);
            return MSL_ERR_UNSUPPORTED;
        }
        smem = sm > smem ? sm : smem;
        DenseParams& p = L.plane[n];
        p.U = q.U; p.u_pitch = q.u_pitch; p.out_he = q.out_he; p.out_clahe = q.out_clahe; p.out_gc = q.out_gc; p.out_lt = q.out_lt;
        p.out_pitch = (size_t)q.rows * q.cols; p.tables = tables;
        p.rows = q.rows; p.cols = q.cols; p.th = q.th; p.tw = q.tw; p.clip = q.clip; p.lut_scale = q.lut_scale;
        if (cl) {
            const size_t tb = dense_tabs_bytes(q.rows, q.cols);
            if (!tabs_ws || (reinterpret_cast<uintptr_t>(tabs_ws) & 15) || tabs_used + tb > tabs_ws_bytes) {
                set_error("enhance_dense: table workspace of %zu bytes (16-byte aligned) needed, %zu given", tabs_used + tb, tabs_ws_bytes);
                return MSL_ERR_WORKSPACE;
            }
            p.ptabs = static_cast<const uint8_t*>(tabs_ws) + tabs_used;
            tabs_used += tb;
        }
        L.first[n] = total;
        total += q.nslices;
        ++n;
    }
    if (n == 0) return MSL_OK;
    for (int i = n; i < 4; ++i) L.first[i] = total;        // unused stacks start behind the grid
    if (cl) {
        ProfScope prof(K_ENH_DENSE_TABLES, stream);
        dense_tables_kernel<<<n, 288, 0, stream>>>(L);
        MSL_LAUNCH_CHECK("dense_tables_kernel");
    }
    ProfScope prof(K_ENH_DENSE, stream);
    if (cl) {
        MSL_CUDA_CHECK(cudaFuncSetAttribute(enhance_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enhance_dense_kernel<true><<<total, kThreads, smem, stream>>>(L);
    } else {
        MSL_CUDA_CHECK(cudaFuncSetAttribute(enhance_dense_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enhance_dense_kernel<false><<<total, kThreads, smem, stream>>>(L);
    }
    MSL_LAUNCH_CHECK("enhance_dense_kernel");
    return MSL_OK;
}

}  // namespace msl
