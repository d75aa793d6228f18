// Dense HE + CLAHE over PNG-oriented uint8 slice stacks (the per-slice stage of msl_enhance_volumes).
//
// Input : the normalised slices U that norm_scatter staged, PNG orientation P[r][c] (c fastest),
//         slice pitch a multiple of 16 bytes.  P[r, c] = G[c, cols-1-r]: P's column index c is the
//         slice ROW a (CLAHE's y, tile height th) and P's row index r mirrors the slice COLUMN
//         b = cols-1-r (CLAHE's x, tile width tw).  The kernel works in P coordinates end to end,
//         so neither the load nor the store transposes anything.
// Output: HE (E3, reference utils/mejora_imagen.py:52-67 == cv2.equalizeHist) and / or CLAHE
//         (E4, :91-117 == LUT_OUT[clahe(LUT_L[u])], OpenCV imgproc/clahe.cpp) slices, densely packed,
//         same orientation.  One CTA per slice; both enhancements share the one shared-memory copy.
//
// Instruction budget matters more than bytes here (the kernel moves 2-3 B per pixel): 128-bit loads,
// 32-bit stores, zero words skipped in the histograms, per-row / per-column interpolation weights
// precomputed once per slice, tile LUTs composed with LUT_L and stored as float so the blend needs no
// int->float conversions, round-half-even through the 1.5*2^23 magic add (no F2I on the XU pipe).
#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

struct DenseParams {
    const uint8_t* U;
    size_t u_pitch;            // bytes between input slices (multiple of 16)
    uint8_t* out_he;           // may be NULL
    uint8_t* out_clahe;        // may be NULL
    uint8_t* out_gc;           // may be NULL  (E5: GC_T[u],      reference utils/mejora_imagen.py:139-151)
    uint8_t* out_lt;           // may be NULL  (E6: LT_T[255][u], reference utils/mejora_imagen.py:166-184)
    size_t out_pitch;          // bytes between output slices (= npx)
    const uint8_t* tables;
    int rows, cols;            // slice orientation (G); P is cols x rows
    int th, tw, clip;
    float lut_scale;
    unsigned magic_w;          // floor(2^32 / rows) + 1 : o / rows == umulhi(o, magic_w) for o < 2^32 / rows
};

struct XY { float w, w1; int o1, o2; };   // blend weights and the two tile offsets (in floats) along one axis

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
    return p;
}

// smem map (bytes): [0,256) lutl | [256,512) lutout | [512,768) gc | [768,1024) lt | [1024,1280) he_lut
//                   | [1280,2304) he_hist u32[256] | [2304,2560) misc | xtab[cols] XY | ytab[rows] XY
//                   | R (64 KB, CLAHE only) | su[npx16]
constexpr int kOffXtab = 2560;

// out[o] = lut[su[o]] for a whole slice: 32-bit smem reads, four byte lookups, 32-bit stores
__device__ __forceinline__ void lut_store(const uint8_t* __restrict__ lut, const uint8_t* __restrict__ su, uint8_t* __restrict__ out, int npx) {
    const uint32_t* su32 = reinterpret_cast<const uint32_t*>(su);
    uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
    const int nw = npx >> 2;
    const uint32_t z4 = (uint32_t)lut[0] * 0x01010101u;
    for (int q = threadIdx.x; q < nw; q += kThreads) {
        uint32_t w = su32[q];
        out32[q] = w ? (uint32_t)lut[w & 0xff] | ((uint32_t)lut[(w >> 8) & 0xff] << 8) |
                           ((uint32_t)lut[(w >> 16) & 0xff] << 16) | ((uint32_t)lut[w >> 24] << 24)
                     : z4;
    }
    for (int o = (nw << 2) + threadIdx.x; o < npx; o += kThreads) out[o] = lut[su[o]];
}

template <bool DO_HE, bool DO_CLAHE>
__global__ void __launch_bounds__(kThreads, 2) enhance_dense_kernel(const DenseParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* lutl = smem;
    uint8_t* lutout = smem + 256;
    uint8_t* gc_lut = smem + 512;
    uint8_t* lt_lut = smem + 768;
    uint8_t* he_lut = smem + 1024;
    unsigned* he_hist = reinterpret_cast<unsigned*>(smem + 1280);
    int* misc = reinterpret_cast<int*>(smem + 2304);       // [0] i0, [1..16] warp scan totals
    const int rows = p.rows, cols = p.cols, npx = rows * cols;
    const int W = rows;                                      // P row length
    XY* xtab = reinterpret_cast<XY*>(smem + kOffXtab);       // indexed by P row r   (slice column b = cols-1-r)
    XY* ytab = xtab + (DO_CLAHE ? cols : 0);                 // indexed by P column c (slice row a = c)
    uint8_t* R = reinterpret_cast<uint8_t*>(ytab + (DO_CLAHE ? rows : 0));
    uint8_t* su = R + (DO_CLAHE ? 65536 : 0);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t s = blockIdx.x;
    const uint8_t* in = p.U + s * p.u_pitch;

    if (tid < 192) reinterpret_cast<uint32_t*>(lutl)[tid] = __ldg(reinterpret_cast<const uint32_t*>(p.tables) + tid);  // lutl + lutout + gc
    else if (tid < 256) reinterpret_cast<uint32_t*>(lt_lut)[tid - 192] = __ldg(reinterpret_cast<const uint32_t*>(p.tables + MSL_TAB_LT + 255 * 256) + (tid - 192));
    if (DO_HE) {
        if (tid < 256) he_hist[tid] = 0;
        if (tid == 0) misc[0] = 256;
    }
    if (DO_CLAHE) {
        uint4* r4 = reinterpret_cast<uint4*>(R);
        for (int q = tid; q < 32768 / 16; q += kThreads) r4[q] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    // ---------------------------------------------------------------- load (+ HE histogram on the fly)
    {
        const int nvec = npx >> 4;
        const uint4* in4 = reinterpret_cast<const uint4*>(in);
        uint4* su4 = reinterpret_cast<uint4*>(su);
        int zeros = 0;
        for (int q = tid; q < nvec; q += kThreads) {
            uint4 v = __ldg(in4 + q);
            su4[q] = v;
            if (DO_HE) {
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w = w4[j];
                    if (w == 0) { zeros += 4; continue; }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t b = (w >> (8 * k)) & 0xff;
                        if (b) atomicAdd(&he_hist[b], 1u); else ++zeros;
                    }
                }
            }
        }
        for (int o = (nvec << 4) + tid; o < npx; o += kThreads) {        // < 16 tail bytes
            uint32_t b = __ldg(in + o);
            su[o] = (uint8_t)b;
            if (DO_HE) { if (b) atomicAdd(&he_hist[b], 1u); else ++zeros; }
        }
        if (DO_HE) {
            zeros = warp_sum(zeros);
            if (lane == 0 && zeros) atomicAdd(&he_hist[0], (unsigned)zeros);
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- HE: CDF -> LUT -> store
    if (DO_HE) {
        int h = 0, c = 0;
        if (tid < 256) {
            h = (int)he_hist[tid];
            c = warp_incl_scan(h, lane);
            if (lane == 31) misc[1 + warp] = c;
            if (h > 0) atomicMin(&misc[0], tid);
        }
        __syncthreads();
        if (tid < 256) {
            for (int w = 0; w < warp; ++w) c += misc[1 + w];
            const int i0 = misc[0];
            const int h0 = (int)he_hist[i0];
            uint8_t o;
            if (h0 == npx) o = (uint8_t)i0;
            else if (tid <= i0) o = 0;
            else o = sat_u8_rn(__fmul_rn((float)(c - h0), __fdiv_rn(255.0f, (float)(npx - h0))));
            he_lut[tid] = o;
        }
        __syncthreads();
        lut_store(he_lut, su, p.out_he + s * p.out_pitch, npx);
    }
    // ---------------------------------------------------------------- GC / LT: table maps of the same smem slice
    // (E1 maps the slice maximum to exactly 255 whenever ptp > 0, so LT's table row is 255; a blank slice is all
    // zeros and LT_T[255][0] == LT_T[0][0] == 0.)
    if (p.out_gc) lut_store(gc_lut, su, p.out_gc + s * p.out_pitch, npx);
    if (p.out_lt) lut_store(lt_lut, su, p.out_lt + s * p.out_pitch, npx);

    if (!DO_CLAHE) return;

    // ---------------------------------------------------------------- CLAHE: tile histograms
    // Tile maps: ty of every P column c (slice row a) and tx of every P row r (slice column b = cols-1-r),
    // written into the not-yet-used tail of the interpolation tables' neighbourhood (misc scratch is too small):
    // they live in the first bytes of the upper half of R (the float LUT area is only needed after the CDFs).
    const int th = p.th, tw = p.tw;
    unsigned* hist = reinterpret_cast<unsigned*>(R);        // 64 tiles x 128 words (two 16-bit bins per word)
    uint8_t* tya = R + 32768;                                // [rows]
    uint8_t* txr = tya + ((rows + 3) & ~3);                  // [cols]
    for (int a = tid; a < rows; a += kThreads) tya[a] = (uint8_t)(a / th);
    for (int r = tid; r < cols; r += kThreads) txr[r] = (uint8_t)((cols - 1 - r) / tw);
    __syncthreads();
    {
        // real pixels: linear over the smem words, L = LUT_L[u]; zero words (background) cost one atomic
        const uint32_t* su32 = reinterpret_cast<const uint32_t*>(su);
        const int nw = npx >> 2;
        for (int q = tid; q < nw; q += kThreads) {
            const unsigned o = 4u * q;
            int r = (int)__umulhi(o, p.magic_w);
            int c = (int)o - r * W;
            const uint32_t w = su32[q];
            if (c + 3 < W) {
                const int tx = txr[r];
                const int t0 = tya[c] * 8 + tx, t3 = tya[c + 3] * 8 + tx;
                if (w == 0 && t0 == t3) { atomicAdd(&hist[t0 * 128], 4u); continue; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int L = lutl[(w >> (8 * k)) & 0xff];
                    const int t = (k == 0) ? t0 : (k == 3 ? t3 : tya[c + k] * 8 + tx);
                    atomicAdd(&hist[t * 128 + (L >> 1)], 1u << ((L & 1) * 16));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int L = lutl[(w >> (8 * k)) & 0xff];
                    atomicAdd(&hist[(tya[c] * 8 + txr[r]) * 128 + (L >> 1)], 1u << ((L & 1) * 16));
                    if (++c == W) { c = 0; ++r; }
                }
            }
        }
        for (int o = (nw << 2) + tid; o < npx; o += kThreads) {
            const int r = o / W, c = o - r * W, L = lutl[su[o]];
            atomicAdd(&hist[(tya[c] * 8 + txr[r]) * 128 + (L >> 1)], 1u << ((L & 1) * 16));
        }
        // BORDER_REFLECT_101 padding (OpenCV pads bottom / right up to 8 tiles): the few padded pixels
        const int prow = th * 8, pcol = tw * 8;
        const int nA = (prow - rows) * pcol;                 // padded slice rows, all padded columns
        const int nB = rows * (pcol - cols);                 // real slice rows, padded columns
        for (int i = tid; i < nA + nB; i += kThreads) {
            int ap, bp;
            if (i < nA) { ap = rows + i / pcol; bp = i % pcol; }
            else { const int j = i - nA, wp = pcol - cols; ap = j / wp; bp = cols + j % wp; }
            const int a = reflect101(ap, rows), b = reflect101(bp, cols);
            const int L = lutl[su[(cols - 1 - b) * W + a]];
            atomicAdd(&hist[((ap / th) * 8 + bp / tw) * 128 + (L >> 1)], 1u << ((L & 1) * 16));
        }
    }
    __syncthreads();
    // ---------------------------------------------------------------- CLAHE: clip + redistribute + CDF -> tile LUTs
    // (OpenCV CLAHE_CalcLut_Body).  One warp per tile, 8 bins per lane.
    for (int t = warp; t < 64; t += kWarps) {
        unsigned* ht = hist + t * 128;
        uint4 w4 = reinterpret_cast<const uint4*>(ht)[lane];
        int hb[8] = {(int)(w4.x & 0xffff), (int)(w4.x >> 16), (int)(w4.y & 0xffff), (int)(w4.y >> 16),
                     (int)(w4.z & 0xffff), (int)(w4.z >> 16), (int)(w4.w & 0xffff), (int)(w4.w >> 16)};
        const int clip = p.clip;
        int clipped = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (hb[k] > clip) { clipped += hb[k] - clip; hb[k] = clip; }
        clipped = warp_sum(clipped);
        const int rb = clipped / 256;
        const int res = clipped - rb * 256;
        // residual: bins 0, step, 2*step, ... (res of them) get one more; walk this lane's 8 bins without dividing per bin
        const int step = res > 0 ? max(256 / res, 1) : 256;
        const int base = lane * 8;
        int kn = (base + step - 1) / step;                   // index of the first multiple of step that is >= base
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
        __syncwarp();
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t o = sat_u8_rn(__fmul_rn((float)(hb[k] + excl), p.lut_scale));
            if (k < 4) lo |= o << (8 * k); else hi |= o << (8 * (k - 4));
        }
        reinterpret_cast<uint2*>(ht)[lane] = make_uint2(lo, hi);      // T[t][L], 256 bytes at the head of the tile's slot
    }
    // interpolation tables (OpenCV CLAHE_Interpolation_Body): blend weights + tile offsets per column / row
    {
        const float inv_tw = __fdiv_rn(1.0f, (float)tw), inv_th = __fdiv_rn(1.0f, (float)th);
        for (int r = tid; r < cols; r += kThreads) {
            const int b = cols - 1 - r;
            float txf = __fsub_rn(__fmul_rn((float)b, inv_tw), 0.5f);
            int t1 = (int)floorf(txf), t2 = t1 + 1;
            XY e;
            e.w = __fsub_rn(txf, (float)t1); e.w1 = __fsub_rn(1.0f, e.w);
            e.o1 = max(t1, 0) * 256; e.o2 = min(t2, 7) * 256;
            xtab[r] = e;
        }
        for (int a = tid; a < rows; a += kThreads) {
            float tyf = __fsub_rn(__fmul_rn((float)a, inv_th), 0.5f);
            int t1 = (int)floorf(tyf), t2 = t1 + 1;
            XY e;
            e.w = __fsub_rn(tyf, (float)t1); e.w1 = __fsub_rn(1.0f, e.w);
            e.o1 = max(t1, 0) * 2048; e.o2 = min(t2, 7) * 2048;
            ytab[a] = e;
        }
    }
    __syncthreads();
    // compose with LUT_L and widen to float: F[t][u] = (float) T[t][LUT_L[u]]  (64 KB overlaying the histograms)
    float* F = reinterpret_cast<float*>(R);
    {
        float fv[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint8_t* T = R + (warp + j * kWarps) * 512;
#pragma unroll
            for (int k = 0; k < 8; ++k) fv[j][k] = (float)T[lutl[lane * 8 + k]];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4* dst = reinterpret_cast<float4*>(F + (warp + j * kWarps) * 256 + lane * 8);
            dst[0] = make_float4(fv[j][0], fv[j][1], fv[j][2], fv[j][3]);
            dst[1] = make_float4(fv[j][4], fv[j][5], fv[j][6], fv[j][7]);
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- CLAHE: bilinear blend + LUT_OUT, in place
    // Consecutive lanes take consecutive pixels (conflict-free table reads); the result byte replaces u in smem.
    for (int o = tid; o < npx; o += kThreads) {
        const int r = (int)__umulhi((unsigned)o, p.magic_w);
        const int c = o - r * W;
        const XY X = xtab[r];
        const XY Y = ytab[c];
        const float* F1 = F + Y.o1 + su[o];
        const float* F2 = F1 + (Y.o2 - Y.o1);
        float top = __fadd_rn(__fmul_rn(F1[X.o1], X.w1), __fmul_rn(F1[X.o2], X.w));
        float bot = __fadd_rn(__fmul_rn(F2[X.o1], X.w1), __fmul_rn(F2[X.o2], X.w));
        float res = __fadd_rn(__fmul_rn(top, Y.w1), __fmul_rn(bot, Y.w));
        // cvRound (half-even) of a value in [0, 255.0001]: the low mantissa bits of res + 1.5 * 2^23
        su[o] = lutout[__float_as_uint(__fadd_rn(res, 12582912.0f)) & 0xffu];
    }
    __syncthreads();
    {
        uint8_t* out = p.out_clahe + s * p.out_pitch;
        const uint32_t* su32 = reinterpret_cast<const uint32_t*>(su);
        uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
        const int nw = npx >> 2;
        for (int q = tid; q < nw; q += kThreads) out32[q] = su32[q];
        for (int o = (nw << 2) + tid; o < npx; o += kThreads) out[o] = su[o];
    }
}

}  // namespace

size_t dense_u_pitch(int npx) { return ((size_t)npx + 15) & ~(size_t)15; }

int launch_enhance_dense(const uint8_t* U, size_t u_pitch, int nslices, int rows, int cols,
                         uint8_t* out_he, uint8_t* out_clahe, uint8_t* out_gc, uint8_t* out_lt, const uint8_t* tables,
                         int th, int tw, int clip, float lut_scale, cudaStream_t stream) {
    if (nslices <= 0 || (!out_he && !out_clahe && !out_gc && !out_lt)) return MSL_OK;
    const int npx = rows * cols;
    DenseParams p;
    p.U = U; p.u_pitch = u_pitch; p.out_he = out_he; p.out_clahe = out_clahe; p.out_gc = out_gc; p.out_lt = out_lt; p.out_pitch = (size_t)npx; p.tables = tables;
    p.rows = rows; p.cols = cols; p.th = th; p.tw = tw; p.clip = clip; p.lut_scale = lut_scale;
    p.magic_w = (unsigned)(0x100000000ull / (unsigned)rows) + 1u;
    const bool he = out_he != nullptr, cl = out_clahe != nullptr;
    size_t smem = kOffXtab + (cl ? (size_t)(rows + cols) * sizeof(XY) + 65536 : 0) + dense_u_pitch(npx);
    if (smem > 227 * 1024 || (u_pitch & 15) || (reinterpret_cast<uintptr_t>(U) & 15) ||
        ((reinterpret_cast<uintptr_t>(out_he) | reinterpret_cast<uintptr_t>(out_clahe) | reinterpret_cast<uintptr_t>(out_gc) |
          reinterpret_cast<uintptr_t>(out_lt)) & 3) || (npx & 3) || rows < 2 ||
        (unsigned long long)npx * (unsigned)rows >= 0x100000000ull) {
        set_error("enhance_dense: unsupported geometry / alignment (%d x %d, %zu B smem)", rows, cols, smem);
        return MSL_ERR_UNSUPPORTED;
    }
    ProfScope prof(K_ENH_DENSE, stream);
#define MSL_LAUNCH_DENSE(HE, CL)                                                                                         \
    do {                                                                                                                 \
        MSL_CUDA_CHECK(cudaFuncSetAttribute(enhance_dense_kernel<HE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        enhance_dense_kernel<HE, CL><<<nslices, kThreads, smem, stream>>>(p);                                            \
    } while (0)
    if (he && cl) MSL_LAUNCH_DENSE(true, true);
    else if (cl) MSL_LAUNCH_DENSE(false, true);
    else if (he) MSL_LAUNCH_DENSE(true, false);
    else MSL_LAUNCH_DENSE(false, false);
#undef MSL_LAUNCH_DENSE
    MSL_LAUNCH_CHECK("enhance_dense_kernel");
    return MSL_OK;
}

}  // namespace msl
