// YOLO label polygons on the device: binary mask slice -> external contours, the arithmetic behind
// anotar_mascaras (scripts/extraer_dataset.py:215-227): ultralytics convert_segment_masks_to_yolo_seg =
// cv2.findContours(mask == value, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) per mask, one label line per contour.
//
// OpenCV's findContours is a raster scan with Suzuki-Abe border following (imgproc/contours.cpp, icvFindNextContour /
// icvFetchContour).  Restated for a GPU as three data-parallel steps and one step that is serial per contour only:
//   1. connected components by lock-free union-find in shared memory (atomicMin on the parent array, root = smallest pixel
//      index): 8-connectivity for the foreground, 4-connectivity for the background (zero frame included), both in ONE
//      parent array because the two pixel sets are disjoint.  The root of a foreground component IS its raster-first pixel,
//      the pixel where OpenCV's scan meets the component and starts the outer border.
//   2. a component is "external" iff the background region to the left of its first pixel is the frame's region (root 0):
//      this is what RETR_EXTERNAL keeps (components inside holes of other components are dropped).
//   3. starts are ranked by a block-wide prefix sum; OpenCV returns the contours in REVERSE raster order of their starts.
//   4. one thread per contour walks the border exactly like icvFetchContour (8 direction codes, first neighbour search from
//      direction 4 clockwise, then counter-clockwise from the arrival direction; CHAIN_APPROX_SIMPLE stores a point only
//      where the direction code changes), first to count the points, then - offsets known - to store them.
// The text formatting of the label lines (round(x / width, 6) ...) stays on the host: a few numbers per contour.
#include <cstring>

#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

constexpr int kCtThreads = 512;

struct CtArgs {
    const uint8_t* masks;      // [n][H][W]
    int n, H, W;
    int value;                 // 0: foreground = any non-zero byte; else foreground = (byte == value)
    int max_contours, max_points;
    uint32_t* counts;          // [n][4]: contours, points, overflow flag (1 contours, 2 points), 0
    uint32_t* contour_len;     // [n][max_contours]
    short2* points;            // [n][max_points] (x, y), contours back to back in OpenCV's order
};

__device__ __forceinline__ int uf_find(const uint32_t* L, int i) {
    int r = (int)L[i];
    while (r != (int)L[r]) r = (int)L[r];
    return r;
}

__device__ __forceinline__ void uf_union(uint32_t* L, int a, int b) {
    for (;;) {
        a = uf_find(L, a); b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          // a > b: hang a under b
        const uint32_t old = atomicMin(&L[a], (uint32_t)b);
        if (old == (uint32_t)a) return;
        a = (int)old;
    }
}

// icvFetchContour for an outer border starting at the component's raster-first pixel p0 (index into the padded image of
// row pitch P).  EMIT = false: returns the number of CHAIN_APPROX_SIMPLE points; EMIT = true: also stores them.
template <bool EMIT>
__device__ int trace_border(const uint8_t* M, int P, int p0, short2* out) {
    const int d[8] = {1, -P + 1, -P, -P - 1, -1, P - 1, P, P + 1};
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int s = 4, s_end = 4, i1;
    do {
        s = (s - 1) & 7;
        i1 = p0 + d[s];
    } while (M[i1] == 0 && s != s_end);
    int x = p0 % P - 1, y = p0 / P - 1, npts = 0;
    if (s == s_end) {                       // single pixel
        if (EMIT) out[0] = make_short2((short)x, (short)y);
        return 1;
    }
    int i3 = p0, i4 = p0, prev_s = s ^ 4;
    for (;;) {
        s_end = s;
        while (s < 15) {
            ++s;
            i4 = i3 + d[s & 7];
            if (M[i4] != 0) break;
        }
        s &= 7;
        if (s != prev_s) {
            if (EMIT) out[npts] = make_short2((short)x, (short)y);
            ++npts;
            prev_s = s;
        }
        x += dx[s]; y += dy[s];
        if (i4 == p0 && i3 == i1) break;
        i3 = i4;
        s = (s + 4) & 7;
    }
    return npts;
}

__global__ void __launch_bounds__(kCtThreads) contours_kernel(const CtArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int H = a.H, W = a.W, P = W + 2, NP = (H + 2) * P;
    uint32_t* L = reinterpret_cast<uint32_t*>(smem);                 // parent array, padded image
    uint8_t* M = smem + (size_t)NP * 4;                              // padded binary image
    __shared__ int scan_w[kCtThreads / 32];
    __shared__ int s_points;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x;
    const uint8_t* src = a.masks + (size_t)s * H * W;
    // ---- padded binary image, parents = self
    for (int i = tid; i < NP; i += kCtThreads) {
        const int y = i / P - 1, x = i % P - 1;
        uint8_t v = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const uint8_t b = __ldg(src + (size_t)y * W + x);
            v = a.value ? (b == a.value) : (b != 0);
        }
        M[i] = v;
        L[i] = (uint32_t)i;
    }
    __syncthreads();
    // ---- unions: foreground with its W, NW, N, NE neighbours (8-connectivity); background with W and N (4-connectivity)
    for (int i = tid; i < NP; i += kCtThreads) {
        const int y = i / P, x = i % P;
        const uint8_t v = M[i];
        if (x > 0 && M[i - 1] == v) uf_union(L, i, i - 1);
        if (y > 0 && M[i - P] == v) uf_union(L, i, i - P);
        if (v && y > 0) {
            if (x > 0 && M[i - P - 1]) uf_union(L, i, i - P - 1);
            if (x < P - 1 && M[i - P + 1]) uf_union(L, i, i - P + 1);
        }
    }
    __syncthreads();
    // ---- contour starts: foreground roots whose left background neighbour belongs to the frame's region (root 0)
    // every thread owns a contiguous run of pixels so that the block scan ranks the starts in raster order
    const int per = (NP + kCtThreads - 1) / kCtThreads;
    const int b0 = tid * per, b1 = min(b0 + per, NP);
    int mine = 0;
    for (int i = b0; i < b1; ++i)
        if (M[i] && (int)L[i] == i && uf_find(L, i - 1) == 0) ++mine;
    int incl = warp_incl_scan(mine, lane);
    if (lane == 31) scan_w[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < kCtThreads / 32; ++w) { const int v = scan_w[w]; if (w < warp) base += v; total += v; }
    base += incl - mine;
    // the slice's row of contour_len doubles as scratch: start pixel of the contour with OpenCV rank k, later its length
    uint32_t* clen = a.contour_len + (size_t)s * a.max_contours;
    const int ncont = min(total, a.max_contours);
    {
        int k = base;
        for (int i = b0; i < b1; ++i)
            if (M[i] && (int)L[i] == i && uf_find(L, i - 1) == 0) {
                const int rank = total - 1 - k;                     // reverse raster order
                if (rank < a.max_contours) clen[rank] = (uint32_t)i;
                ++k;
            }
    }
    __syncthreads();
    // ---- pass 1: points per contour (one thread per contour); the start pixel moves to a register
    uint32_t overflow = total > a.max_contours ? 1u : 0u;
    if (tid == 0) s_points = 0;
    __syncthreads();
    short2* pts = a.points + (size_t)s * a.max_points;
    for (int c0 = 0; c0 < ncont; c0 += kCtThreads) {
        const int c = c0 + tid;
        int start = 0, np = 0;
        if (c < ncont) { start = (int)clen[c]; np = trace_border<false>(M, P, start, nullptr); }
        // exclusive offsets of this batch of contours (in contour order)
        int inc2 = warp_incl_scan(np, lane);
        __syncthreads();
        if (lane == 31) scan_w[warp] = inc2;
        __syncthreads();
        int wb = 0, tot2 = 0;
        for (int w = 0; w < kCtThreads / 32; ++w) { const int v = scan_w[w]; if (w < warp) wb += v; tot2 += v; }
        const int off = s_points + wb + inc2 - np;
        if (c < ncont) {
            if (off + np <= a.max_points) {
                trace_border<true>(M, P, start, pts + off);
                clen[c] = (uint32_t)np;
            } else {
                clen[c] = 0;
                overflow |= 2u;
            }
        }
        __syncthreads();
        if (tid == 0) s_points += tot2;
        __syncthreads();
    }
    if (overflow) atomicOr(&a.counts[4 * (size_t)s + 2], overflow);      // bit 0: more contours than max_contours, bit 1: points
    if (tid == 0) {
        a.counts[4 * (size_t)s + 0] = (uint32_t)ncont;
        a.counts[4 * (size_t)s + 1] = (uint32_t)min(s_points, a.max_points);
        a.counts[4 * (size_t)s + 3] = (uint32_t)total;
    }
}

}  // namespace

size_t contours_smem_bytes(int H, int W) { return (size_t)(H + 2) * (W + 2) * 5 + 16; }

int launch_contours(const uint8_t* masks, int n, int H, int W, int value, int max_contours, int max_points, uint32_t* counts,
                    uint32_t* contour_len, short* points, cudaStream_t stream) {
    if (n <= 0) return MSL_OK;
    const size_t smem = contours_smem_bytes(H, W);
    if (smem > 227 * 1024 || H > 32000 || W > 32000) {
        set_error("contours: a %d x %d mask needs %zu bytes of shared memory (limit 227 KB)", H, W, smem);
        return MSL_ERR_UNSUPPORTED;
    }
    CtArgs a;
    a.masks = masks; a.n = n; a.H = H; a.W = W; a.value = value; a.max_contours = max_contours; a.max_points = max_points;
    a.counts = counts; a.contour_len = contour_len; a.points = reinterpret_cast<short2*>(points);
    MSL_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)n * 16, stream));
    ProfScope prof(K_CONTOURS, stream);
    MSL_CUDA_CHECK(cudaFuncSetAttribute(contours_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    contours_kernel<<<n, kCtThreads, smem, stream>>>(a);
    MSL_LAUNCH_CHECK("contours_kernel");
    return MSL_OK;
}

}  // namespace msl
