// Slice-mode enhancement: one CTA per 2-D slice, the normalised uint8 slice lives in shared
// memory from the first global read to the final store.
//
// Reference rows: E1 normalizar_a_uint8 (utils/utils.py:396-406),
// E2 slice gather + PNG orientation (utils/Paciente.py:230-246, scripts/extraer_dataset.py:192),
// E3 HE (utils/mejora_imagen.py:52-67 == cv2.equalizeHist), E4 CLAHE (:91-117 == OpenCV
// clahe.cpp between the two Lab tables), E5 GC (:139-151), E6 LT (:166-184), E8 imsave
// normalisation + gray colormap (scripts/extraer_dataset.py:192,197; restated, unpinned).
//
// Phases inside the CTA:
//   A  global -> smem: per-slice min/max (float32), then u = trunc(255*((f-min)/ptp)) into su[]
//   B  per-enhancement table(s): 256-bin histogram -> CDF -> LUT (HE); 8x8 tile histograms ->
//      clip/redistribute -> CDF -> tile LUTs (CLAHE); table row copy (GC / LT)
//   C  su[] <- G in place (LUT map, or CLAHE bilinear blend of four tile LUTs); min/max of G when
//      a PNG layout asks for matplotlib's second normalisation
//   D  smem -> global in the requested layout, 32-bit stores when alignment allows
#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

// shared memory map (bytes)
constexpr int kOffTabs = 0;        // lutl[256] lutout[256] ptab[256] cm[256]
constexpr int kOffMisc = 1024;     // 512 B scratch
constexpr int kOffHist = 1536;     // HE: 256 x u32 ; CLAHE: 64 tiles x 128 x u32 (two u16 bins per word)
constexpr int kHistBytesHE = 1024;
constexpr int kHistBytesCLAHE = 64 * 512;

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
    return p;
}

struct Misc {
    float red_min[kWarps];
    float red_max[kWarps];
    int   red_i[kWarps];
    float mn, mx;
    int   i0;
    int   umax;
    int   gmin, gmax;
};

__device__ __forceinline__ void block_minmax(float& mn, float& mx, Misc* m) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) { m->red_min[w] = mn; m->red_max[w] = mx; }
    __syncthreads();
    if (w == 0) {
        float a = lane < kWarps ? m->red_min[lane] : m->red_min[0];
        float b = lane < kWarps ? m->red_max[lane] : m->red_max[0];
        a = warp_min(a);
        b = warp_max(b);
        if (lane == 0) { m->mn = a; m->mx = b; }
    }
    __syncthreads();
    mn = m->mn;
    mx = m->mx;
}

template <typename InT>
__global__ void __launch_bounds__(kThreads) enhance_slices_kernel(const EnhParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* lutl = smem + kOffTabs;
    uint8_t* lutout = lutl + 256;
    uint8_t* ptab = lutl + 512;
    uint8_t* cm = lutl + 768;
    Misc* misc = reinterpret_cast<Misc*>(smem + kOffMisc);
    unsigned* hist = reinterpret_cast<unsigned*>(smem + kOffHist);
    uint8_t* su = smem + kOffHist + p.hist_bytes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x;
    int v, i;
    if (p.vol_of_slice) { v = p.vol_of_slice[s]; i = p.idx_of_slice[s]; }
    else { v = s / p.n_plane; i = s - v * p.n_plane; }
    if (v < 0 || v >= p.nvol || i < 0 || i >= p.n_plane) return;

    const InT* in = reinterpret_cast<const InT*>(p.in) + (long long)v * p.vol_stride + (long long)i * p.idx_stride + p.base0;
    uint8_t* out = p.out + (size_t)s * p.out_pitch;
    const int rows = p.rows, cols = p.cols, npx = rows * cols;
    const long long sa = p.sa, sb = p.sb;
    const bool fast_a = (sa < 0 ? -sa : sa) < (sb < 0 ? -sb : sb);
    const int nline = fast_a ? cols : rows, len = fast_a ? rows : cols;
    const long long s_line = fast_a ? sb : sa, s_elem = fast_a ? sa : sb;
    const bool png = p.layout >= MSL_OUT_PNG_GRAY;

    // constant tables -> smem (1 KB)
    if (tid < 256) {
        reinterpret_cast<uint32_t*>(lutl)[tid] = __ldg(reinterpret_cast<const uint32_t*>(p.tables) + tid);
    }

    // ------------------------------------------------------------------ phase A
    float mn = 0.f, mx = 0.f;
    if (sizeof(InT) == 4) {
        mn = INFINITY; mx = -INFINITY;
        for (int l = warp; l < nline; l += kWarps) {
            const InT* row = in + (long long)l * s_line;
            for (int e = lane; e < len; e += 32) {
                float f = load_as_float(row + (long long)e * s_elem);
                mn = fminf(mn, f);
                mx = fmaxf(mx, f);
            }
        }
        block_minmax(mn, mx, misc);
    }

    if (p.mejora == MSL_MEJORA_NONE && sizeof(InT) == 4 && png) {
        // imsave of the raw float slice: float64 normalisation straight from global memory
        // (scripts/extraer_dataset.py:192 with mejora=None; matplotlib Normalize on float64 input).
        __syncthreads();
        const double vmin = (double)mn, vmax = (double)mx, den = vmax - vmin;
        const int W = rows;   // PNG orientation: (cols, rows)
        for (int o = tid; o < npx; o += kThreads) {
            int r = o / W, c = o - r * W;
            int a = c, b = cols - 1 - r;
            uint8_t g = 0;
            if (vmin != vmax) {
                double t = ((double)load_as_float(in + (long long)a * sa + (long long)b * sb) - vmin) / den;
                t = t * 256.0;
                if (t == 256.0) t = 255.0;
                g = cm[(int)t];
            }
            if (p.layout == MSL_OUT_PNG_RGBA)
                reinterpret_cast<uint32_t*>(out)[o] = 0xff000000u | (g * 0x010101u);
            else
                out[o] = g;
        }
        return;
    }

    int umax = 0;
    if (sizeof(InT) == 4) {
        const float ptp = __fsub_rn(mx, mn);
        for (int l = warp; l < nline; l += kWarps) {
            const InT* row = in + (long long)l * s_line;
            for (int e = lane; e < len; e += 32) {
                float f = load_as_float(row + (long long)e * s_elem);
                int a = fast_a ? e : l, b = fast_a ? l : e;
                su[a * cols + b] = normalise_px(f, mn, ptp);
            }
        }
        umax = ptp > 0.f ? 255 : 0;
    } else {
        int m = 0;
        for (int l = warp; l < nline; l += kWarps) {
            const InT* row = in + (long long)l * s_line;
            for (int e = lane; e < len; e += 32) {
                int u = (int)__ldg(reinterpret_cast<const uint8_t*>(row + (long long)e * s_elem));
                int a = fast_a ? e : l, b = fast_a ? l : e;
                su[a * cols + b] = (uint8_t)u;
                m = max(m, u);
            }
        }
        if (p.mejora == MSL_MEJORA_LT) {
            float fm = (float)m, dummy = fm;
            block_minmax(dummy, fm, misc);
            umax = (int)fm;
        }
    }

    // ------------------------------------------------------------------ phase B
    if (p.mejora == MSL_MEJORA_HE) {
        if (tid < 256) hist[tid] = 0;
        if (tid == 0) misc->i0 = 256;
        __syncthreads();
        // smem-privatised histogram; zeros (the skull-stripped background, ~3/4 of the pixels) are
        // counted with ballot/popc instead of hammering one bank with atomics.
        int zeros = 0;
        const int npx4 = (npx + 3) >> 2;
        const uint32_t* su4 = reinterpret_cast<const uint32_t*>(su);
        for (int q = tid; q < ((npx4 + 31) & ~31); q += kThreads) {
            uint32_t w4 = 0; int nvalid = 0;
            if (q < npx4) { w4 = su4[q]; nvalid = min(4, npx - 4 * q); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int val = (w4 >> (8 * k)) & 0xff;
                bool ok = k < nvalid;
                bool z = ok && val == 0;
                unsigned bz = __ballot_sync(FULL, z);
                zeros += __popc(bz);
                if (ok && !z) atomicAdd(&hist[val], 1u);
            }
        }
        if (lane == 0 && zeros) atomicAdd(&hist[0], (unsigned)zeros);
        __syncthreads();
        // CDF -> LUT (OpenCV equalizeHist)
        int h = 0, c = 0;
        if (tid < 256) {
            h = (int)hist[tid];
            c = warp_incl_scan(h, lane);
            if (lane == 31) misc->red_i[warp] = c;
            if (h > 0) atomicMin(&misc->i0, tid);
        }
        __syncthreads();
        if (tid < 256) {
            for (int w = 0; w < warp; ++w) c += misc->red_i[w];
            const int i0 = misc->i0;
            const int h0 = (int)hist[i0];
            uint8_t o;
            if (h0 == npx) o = (uint8_t)i0;
            else if (tid <= i0) o = 0;
            else {
                float scale = __fdiv_rn(255.0f, (float)(npx - h0));
                o = sat_u8_rn(__fmul_rn((float)(c - h0), scale));
            }
            ptab[tid] = o;
        }
        __syncthreads();
    } else if (p.mejora == MSL_MEJORA_GC) {
        __syncthreads();
        if (tid < 64) reinterpret_cast<uint32_t*>(ptab)[tid] = __ldg(reinterpret_cast<const uint32_t*>(p.tables + MSL_TAB_GC) + tid);
        __syncthreads();
    } else if (p.mejora == MSL_MEJORA_LT) {
        __syncthreads();
        if (tid < 64) reinterpret_cast<uint32_t*>(ptab)[tid] = __ldg(reinterpret_cast<const uint32_t*>(p.tables + MSL_TAB_LT + umax * 256) + tid);
        __syncthreads();
    } else if (p.mejora == MSL_MEJORA_CLAHE) {
        for (int q = tid; q < 64 * 128; q += kThreads) hist[q] = 0;
        __syncthreads();
        const int th = p.cl_th, tw = p.cl_tw, area = th * tw;
        // one warp per tile: histogram of L = LUT_L[u] over the REFLECT_101-padded tile
        for (int t = warp; t < 64; t += kWarps) {
            const int ty = t >> 3, tx = t & 7;
            unsigned* ht = hist + t * 128;
            int zeros = 0;
            for (int l0 = 0; l0 < area; l0 += 32) {
                int l = l0 + lane;
                bool ok = l < area;
                int L = 0;
                if (ok) {
                    int yy = l / tw, xx = l - yy * tw;
                    int ya = reflect101(ty * th + yy, rows), xb = reflect101(tx * tw + xx, cols);
                    L = lutl[su[ya * cols + xb]];
                }
                bool z = ok && L == 0;
                zeros += __popc(__ballot_sync(FULL, z));
                if (ok && !z) atomicAdd(&ht[L >> 1], 1u << ((L & 1) * 16));
            }
            __syncwarp();
            // clip + redistribute + CDF -> tile LUT (OpenCV CLAHE_CalcLut_Body)
            uint4 w4 = reinterpret_cast<const uint4*>(ht)[lane];
            int hb[8] = {(int)(w4.x & 0xffff), (int)(w4.x >> 16), (int)(w4.y & 0xffff), (int)(w4.y >> 16),
                         (int)(w4.z & 0xffff), (int)(w4.z >> 16), (int)(w4.w & 0xffff), (int)(w4.w >> 16)};
            if (lane == 0) hb[0] += zeros;
            const int clip = p.cl_clip;
            int clipped = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (hb[k] > clip) { clipped += hb[k] - clip; hb[k] = clip; }
            }
            clipped = warp_sum(clipped);
            const int rb = clipped / 256;
            int res = clipped - rb * 256;
            const int step = res > 0 ? max(256 / res, 1) : 1;
            int run = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int bin = lane * 8 + k;
                hb[k] += rb;
                if (res > 0 && (bin % step) == 0 && (bin / step) < res) hb[k] += 1;
                run += hb[k];
                hb[k] = run;
            }
            int excl = warp_incl_scan(run, lane) - run;
            __syncwarp();
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t o = sat_u8_rn(__fmul_rn((float)(hb[k] + excl), p.cl_lut_scale));
                if (k < 4) lo |= o << (8 * k); else hi |= o << (8 * (k - 4));
            }
            reinterpret_cast<uint2*>(ht)[lane] = make_uint2(lo, hi);   // tile LUT overlays its histogram
        }
        __syncthreads();
    } else {
        __syncthreads();
    }

    // ------------------------------------------------------------------ phase C: su <- G
    int gmin = 255, gmax = 0;
    if (p.mejora == MSL_MEJORA_CLAHE) {
        const int th = p.cl_th, tw = p.cl_tw;
        const float inv_tw = __fdiv_rn(1.0f, (float)tw), inv_th = __fdiv_rn(1.0f, (float)th);
        const uint8_t* luts = reinterpret_cast<const uint8_t*>(hist);
        for (int a = warp; a < rows; a += kWarps) {
            float tyf = __fsub_rn(__fmul_rn((float)a, inv_th), 0.5f);
            int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
            float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
            ty1 = max(ty1, 0); ty2 = min(ty2, 7);
            for (int b = lane; b < cols; b += 32) {
                float txf = __fsub_rn(__fmul_rn((float)b, inv_tw), 0.5f);
                int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
                float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
                tx1 = max(tx1, 0); tx2 = min(tx2, 7);
                int val = lutl[su[a * cols + b]];
                float l11 = (float)luts[(ty1 * 8 + tx1) * 512 + val], l12 = (float)luts[(ty1 * 8 + tx2) * 512 + val];
                float l21 = (float)luts[(ty2 * 8 + tx1) * 512 + val], l22 = (float)luts[(ty2 * 8 + tx2) * 512 + val];
                float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
                float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
                float r = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
                int g = lutout[sat_u8_rn(r)];
                su[a * cols + b] = (uint8_t)g;
                gmin = min(gmin, g); gmax = max(gmax, g);
            }
        }
    } else if (p.mejora != MSL_MEJORA_NONE) {
        uint32_t* su4 = reinterpret_cast<uint32_t*>(su);
        const int npx4 = (npx + 3) >> 2;     // su is padded to a multiple of 4 bytes
        for (int q = tid; q < npx4; q += kThreads) {
            uint32_t w4 = su4[q];
            int g0 = ptab[w4 & 0xff], g1 = ptab[(w4 >> 8) & 0xff], g2 = ptab[(w4 >> 16) & 0xff], g3 = ptab[w4 >> 24];
            su4[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
            if (png) {
                int nvalid = min(4, npx - 4 * q);
                gmin = min(gmin, g0); gmax = max(gmax, g0);
                if (nvalid > 1) { gmin = min(gmin, g1); gmax = max(gmax, g1); }
                if (nvalid > 2) { gmin = min(gmin, g2); gmax = max(gmax, g2); }
                if (nvalid > 3) { gmin = min(gmin, g3); gmax = max(gmax, g3); }
            }
        }
    } else if (png) {
        for (int q = tid; q < npx; q += kThreads) { int g = su[q]; gmin = min(gmin, g); gmax = max(gmax, g); }
    }
    float fgmin = 0.f, fgden = 0.f;
    if (png) {
        float a = (float)gmin, b = (float)gmax;
        __syncthreads();
        block_minmax(a, b, misc);
        fgmin = a;
        fgden = __fsub_rn(b, a);
    } else {
        __syncthreads();
    }

    // ------------------------------------------------------------------ phase D: store
    const bool layoutG = p.layout == MSL_OUT_G;
    const int W = layoutG ? cols : rows;
    auto fetch = [&](int q, int rem) -> uint32_t {
        // (q, rem) = divmod(output index, W)
        int a = layoutG ? q : rem, b = layoutG ? rem : cols - 1 - q;
        uint32_t g = su[a * cols + b];
        if (png) {
            // matplotlib Normalize + Colormap on integer input: float32 (g-vmin)/(vmax-vmin)*256
            if (fgden == 0.f) g = 0;
            else {
                float t = __fmul_rn(__fdiv_rn(__fsub_rn((float)g, fgmin), fgden), 256.0f);
                if (t == 256.0f) t = 255.0f;
                g = (uint32_t)(int)t;
            }
            g = cm[g];
        }
        return g;
    };
    if (p.layout == MSL_OUT_PNG_RGBA) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
        for (int o = tid; o < npx; o += kThreads) {
            int q = o / W, rem = o - q * W;
            o32[o] = 0xff000000u | (fetch(q, rem) * 0x010101u);
        }
    } else {
        const bool al4 = ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
        for (int o4 = tid * 4; o4 < npx; o4 += kThreads * 4) {
            int q = o4 / W, rem = o4 - q * W;
            uint32_t packed = 0;
            const int n = min(4, npx - o4);
            for (int k = 0; k < n; ++k) {
                packed |= fetch(q, rem) << (8 * k);
                if (++rem == W) { rem = 0; ++q; }
            }
            if (al4 && n == 4) *reinterpret_cast<uint32_t*>(out + o4) = packed;
            else for (int k = 0; k < n; ++k) out[o4 + k] = (uint8_t)(packed >> (8 * k));
        }
    }
}

}  // namespace

size_t enhance_slices_smem_bytes(int mejora, int npx) {
    size_t hist = mejora == MSL_MEJORA_CLAHE ? kHistBytesCLAHE : kHistBytesHE;
    return (size_t)kOffHist + hist + (((size_t)npx + 15) & ~(size_t)15);
}

// cv2.COLOR_BGR2GRAY, 8-bit path: descale(B * 3735 + G * 19235 + R * 9798, 15) (OpenCV color_yuv / color_rgb fixed point)
__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, size_t npx, uint8_t* __restrict__ gray) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15);
    }
}

int launch_bgr_to_gray(const uint8_t* bgr, size_t npx, uint8_t* gray, cudaStream_t stream) {
    ProfScope prof(K_BGR2GRAY, stream);
    const unsigned blocks = (unsigned)((npx + 255) / 256 < 148 * 8 ? (npx + 255) / 256 : 148 * 8);
    bgr_to_gray_kernel<<<blocks, 256, 0, stream>>>(bgr, npx, gray);
    MSL_LAUNCH_CHECK("bgr_to_gray_kernel");
    return MSL_OK;
}

int launch_enhance_slices(EnhParams p, int dtype, int nslices, cudaStream_t stream) {
    const int npx = p.rows * p.cols;
    p.hist_bytes = p.mejora == MSL_MEJORA_CLAHE ? kHistBytesCLAHE : kHistBytesHE;
    const size_t smem = enhance_slices_smem_bytes(p.mejora, npx);
    if (smem > 227 * 1024) {
        set_error("slice of %d x %d pixels needs %zu bytes of shared memory (> 227 KB)", p.rows, p.cols, smem);
        return MSL_ERR_UNSUPPORTED;
    }
    if (nslices <= 0) return MSL_OK;
    ProfScope prof((dtype == MSL_F32 ? K_ENH_F32 : K_ENH_U8) + p.mejora, stream);
    if (dtype == MSL_F32) {
        MSL_CUDA_CHECK(cudaFuncSetAttribute(enhance_slices_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enhance_slices_kernel<float><<<nslices, kThreads, smem, stream>>>(p);
    } else {
        MSL_CUDA_CHECK(cudaFuncSetAttribute(enhance_slices_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enhance_slices_kernel<uint8_t><<<nslices, kThreads, smem, stream>>>(p);
    }
    MSL_LAUNCH_CHECK("enhance_slices_kernel");
    return MSL_OK;
}

}  // namespace msl
