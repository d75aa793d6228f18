// extern "C" entry points of libmslesseg.so (declared in include/mslesseg.h): argument validation,
// geometry (plane -> strides, OpenCV's CLAHE tile geometry) and the launch sequences.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

#define MSL_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            msl::set_error(__VA_ARGS__);            \
            return MSL_ERR_ARG;                     \
        }                                           \
    } while (0)

inline int n_plane_of(int plano, int X, int Y, int Z) { return plano == MSL_AXIAL ? Z : (plano == MSL_CORONAL ? Y : X); }

// OpenCV CLAHE_Impl::apply geometry for clipLimit 2.0 and an 8x8 grid (reference
// utils/mejora_imagen.py:86,104).
void clahe_geometry(int rows, int cols, EnhParams& p) {
    int prow = rows, pcol = cols;
    if (!(cols % 8 == 0 && rows % 8 == 0)) {
        prow = rows + (8 - rows % 8);
        pcol = cols + (8 - cols % 8);
    }
    p.cl_th = prow / 8;
    p.cl_tw = pcol / 8;
    const int area = p.cl_th * p.cl_tw;
    int clip = (int)(2.0 * area / 256);
    p.cl_clip = clip > 1 ? clip : 1;
    p.cl_lut_scale = 255.0f / (float)area;
}

int check_enhance_combo(int mejora, int dtype, int layout) {
    MSL_REQUIRE(mejora >= MSL_MEJORA_NONE && mejora <= MSL_MEJORA_LT, "mejora %d no reconocida", mejora);
    MSL_REQUIRE(dtype == MSL_F32 || dtype == MSL_U8, "dtype %d not in {MSL_F32, MSL_U8}", dtype);
    MSL_REQUIRE(layout >= MSL_OUT_G && layout <= MSL_OUT_PNG_RGBA, "layout %d not in MSL_OUT_*", layout);
    return MSL_OK;
}

int check_out(const uint8_t* out, size_t pitch, int npx, int layout) {
    MSL_REQUIRE(out != nullptr, "out is NULL");
    const size_t need = layout == MSL_OUT_PNG_RGBA ? (size_t)npx * 4 : (size_t)npx;
    MSL_REQUIRE(pitch >= need, "output pitch %zu smaller than a slice (%zu bytes)", pitch, need);
    if (layout == MSL_OUT_PNG_RGBA)
        MSL_REQUIRE(pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0, "RGBA output must be 4-byte aligned");
    return MSL_OK;
}

}  // namespace
}  // namespace msl

using namespace msl;

extern "C" {

int msl_version(void) { return MSL_ABI_VERSION; }

const char* msl_last_error(void) { return g_err; }

// bytes of the three staged uint8 slice stacks for ONE volume (16-byte slice pitch, 256-byte stack pitch)
static size_t u_stack_bytes_per_volume(int X, int Y, int Z) {
    const size_t ax = (size_t)Z * dense_u_pitch(X * Y), co = (size_t)Y * dense_u_pitch(X * Z), sa = (size_t)X * dense_u_pitch(Y * Z);
    return ((ax + 255) & ~(size_t)255) + ((co + 255) & ~(size_t)255) + ((sa + 255) & ~(size_t)255);
}

// per-plane table blocks of the dense kernel (three planes)
static size_t dense_tabs_total(int X, int Y, int Z) {
    return ((dense_tabs_bytes(X, Y) + dense_tabs_bytes(X, Z) + dense_tabs_bytes(Y, Z)) + 255) & ~(size_t)255;
}

static int enhance_chunk_volumes(void) {
    const char* e = getenv("MSL_VOLUME_CHUNK");
    int c = e ? atoi(e) : 32;
    return c < 1 ? 1 : c;
}

size_t msl_workspace_bytes(int op, int nvol, int X, int Y, int Z) {
    if (nvol <= 0 || X <= 0 || Y <= 0 || Z <= 0) return 0;
    const size_t N = (size_t)X * Y * Z;
    switch (op) {
        case MSL_WS_ENHANCE_VOLUMES: {
            size_t stats = (((size_t)nvol * (X + Y + Z) * 2 * sizeof(unsigned)) + 255) & ~(size_t)255;
            int chunk = enhance_chunk_volumes();
            if (chunk > nvol) chunk = nvol;
            (void)N;
            return stats + (size_t)chunk * u_stack_bytes_per_volume(X, Y, Z) + dense_tabs_total(X, Y, Z);
        }
        case MSL_WS_RECON: {
            int m = X > Y ? X : Y;
            if (Z > m) m = Z;
            return (size_t)nvol * (m + 2) * sizeof(int32_t);      // inverse slice map + first/last present index
        }
        default:
            return 0;
    }
}

int msl_lesion_slices(const void* gt, int dtype, int nvol, int X, int Y, int Z,
                      uint8_t* any_ax, uint8_t* any_co, uint8_t* any_sa, msl_stream_t stream) {
    MSL_REQUIRE(gt && any_ax && any_co && any_sa, "NULL pointer");
    MSL_REQUIRE(dtype == MSL_F32 || dtype == MSL_U8, "dtype %d not in {MSL_F32, MSL_U8}", dtype);
    MSL_REQUIRE(nvol > 0 && X > 0 && Y > 0 && Z > 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    return launch_lesion_flags(gt, dtype, nvol, X, Y, Z, any_ax, any_co, any_sa, (cudaStream_t)stream);
}

int msl_selftest_norm_division(const float* g, const float* p, size_t n, unsigned long long* out4, msl_stream_t stream_) {
    if ((!g || !p) && n) { set_error("msl_selftest_norm_division: NULL operands"); return MSL_ERR_ARG; }
    if (!out4) { set_error("msl_selftest_norm_division: NULL result"); return MSL_ERR_ARG; }
    return launch_selftest_norm_division(g, p, n, out4, reinterpret_cast<cudaStream_t>(stream_));
}

int msl_slice_ranges(const float* vol, int nvol, int X, int Y, int Z, float* ranges, msl_stream_t stream_) {
    MSL_REQUIRE(vol && ranges, "NULL pointer");
    MSL_REQUIRE(nvol > 0 && X > 0 && Y > 0 && Z > 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    if (X > 256) { set_error("msl_slice_ranges supports X <= 256 (got %d)", X); return MSL_ERR_UNSUPPORTED; }
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned* stats = reinterpret_cast<unsigned*>(ranges);
    const size_t n = (size_t)nvol * (X + Y + Z);
    int rc = launch_init_stats(stats, n, stream);
    if (rc) return rc;
    rc = launch_plane_stats_f32(vol, nvol, X, Y, Z, stats, stream);
    if (rc) return rc;
    return launch_stats_keys_to_float(stats, n * 2, stream);
}

int msl_enhance_slices(const void* vol, int dtype, int nvol, int X, int Y, int Z, int mejora, int plano,
                       const int32_t* vol_of_slice, const int32_t* idx_of_slice, int nslices,
                       uint8_t* out, size_t slice_pitch_bytes, int layout, const uint8_t* tables, msl_stream_t stream) {
    MSL_REQUIRE(vol && tables, "NULL pointer");
    MSL_REQUIRE(nvol > 0 && X > 0 && Y > 0 && Z > 0 && nslices >= 0, "non-positive size");
    MSL_REQUIRE(plano >= MSL_AXIAL && plano <= MSL_SAGITAL, "Plano %d no válido.", plano);
    int rc = check_enhance_combo(mejora, dtype, layout);
    if (rc) return rc;
    MSL_REQUIRE((vol_of_slice == nullptr) == (idx_of_slice == nullptr), "vol_of_slice and idx_of_slice must both be given or both be NULL");
    EnhParams p;
    memset(&p, 0, sizeof(p));
    const long long N = (long long)X * Y * Z;
    p.in = vol;
    p.vol_stride = N;
    p.nvol = nvol;
    p.n_plane = n_plane_of(plano, X, Y, Z);
    if (plano == MSL_AXIAL) { p.rows = X; p.cols = Y; p.sa = 1; p.sb = X; p.idx_stride = (long long)X * Y; }
    else if (plano == MSL_CORONAL) { p.rows = X; p.cols = Z; p.sa = 1; p.sb = (long long)X * Y; p.idx_stride = X; }
    else { p.rows = Y; p.cols = Z; p.sa = X; p.sb = (long long)X * Y; p.idx_stride = 1; }
    if (!vol_of_slice)
        MSL_REQUIRE((long long)nslices == (long long)nvol * p.n_plane, "dense mode needs nslices == nvol * n_plane (%lld), got %d",
                    (long long)nvol * p.n_plane, nslices);
    if (nslices == 0) return MSL_OK;
    rc = check_out(out, slice_pitch_bytes, p.rows * p.cols, layout);
    if (rc) return rc;
    p.vol_of_slice = vol_of_slice; p.idx_of_slice = idx_of_slice;
    p.out = out; p.out_pitch = slice_pitch_bytes; p.layout = layout; p.mejora = mejora; p.tables = tables;
    clahe_geometry(p.rows, p.cols, p);
    return launch_enhance_slices(p, dtype, nslices, (cudaStream_t)stream);
}

int msl_enhance_images(const void* imgs, int dtype, int nimg, int rows, int cols, size_t img_pitch_elems,
                       int mejora, uint8_t* out, size_t out_pitch_bytes, int layout, const uint8_t* tables, msl_stream_t stream) {
    MSL_REQUIRE(imgs && tables, "NULL pointer");
    MSL_REQUIRE(nimg >= 0 && rows > 0 && cols > 0, "non-positive size");
    MSL_REQUIRE(img_pitch_elems >= (size_t)rows * cols, "image pitch smaller than an image");
    int rc = check_enhance_combo(mejora, dtype, layout);
    if (rc) return rc;
    if (nimg == 0) return MSL_OK;
    rc = check_out(out, out_pitch_bytes, rows * cols, layout);
    if (rc) return rc;
    EnhParams p;
    memset(&p, 0, sizeof(p));
    p.in = imgs; p.vol_stride = (long long)img_pitch_elems; p.idx_stride = 0; p.base0 = 0;
    p.sa = cols; p.sb = 1; p.rows = rows; p.cols = cols; p.nvol = nimg; p.n_plane = 1;
    p.out = out; p.out_pitch = out_pitch_bytes; p.layout = layout; p.mejora = mejora; p.tables = tables;
    clahe_geometry(rows, cols, p);
    return launch_enhance_slices(p, dtype, nimg, (cudaStream_t)stream);
}

int msl_enhance_volumes(const float* vol, int nvol, int X, int Y, int Z, uint8_t* const* outs, const uint8_t* tables,
                        void* ws, size_t ws_bytes, msl_stream_t stream_) {
    MSL_REQUIRE(vol && outs && tables, "NULL pointer");
    MSL_REQUIRE(nvol > 0 && X > 0 && Y > 0 && Z > 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    if (X > 256) { set_error("msl_enhance_volumes supports X <= 256 (got %d)", X); return MSL_ERR_UNSUPPORTED; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t need = msl_workspace_bytes(MSL_WS_ENHANCE_VOLUMES, nvol, X, Y, Z);
    if (!ws || ws_bytes < need) { set_error("workspace of %zu bytes needed, %zu given", need, ws_bytes); return MSL_ERR_WORKSPACE; }
    bool any = false;
    for (int k = 0; k < 12; ++k) any |= outs[k] != nullptr;
    if (!any) return MSL_OK;

    const size_t N = (size_t)X * Y * Z;
    const size_t nsl = (size_t)nvol * (X + Y + Z);
    unsigned* stats = reinterpret_cast<unsigned*>(ws);
    const size_t stats_bytes = ((nsl * 2 * sizeof(unsigned)) + 255) & ~(size_t)255;
    int chunk = enhance_chunk_volumes();
    if (chunk > nvol) chunk = nvol;
    const int n_p[3] = {Z, Y, X};
    const int rows_p[3] = {X, X, Y}, cols_p[3] = {Y, Z, Z};
    // staged normalised stacks, PNG orientation, slice pitch padded to 16 bytes: [chunk][n_p][u_pitch]
    size_t upitch[3];
    uint8_t* U[3];
    uint8_t* dtabs = nullptr;                       // per-plane tables of the dense kernel
    const size_t dtabs_bytes = dense_tabs_total(X, Y, Z);
    {
        uint8_t* cur = reinterpret_cast<uint8_t*>(ws) + stats_bytes;
        for (int pl = 0; pl < 3; ++pl) {
            upitch[pl] = dense_u_pitch(rows_p[pl] * cols_p[pl]);
            U[pl] = cur;
            cur += (((size_t)n_p[pl] * upitch[pl] + 255) & ~(size_t)255) * (size_t)chunk;
        }
        dtabs = cur;
    }

    int rc = launch_init_stats(stats, nsl, stream);
    if (rc) return rc;
    for (int v0 = 0; v0 < nvol; v0 += chunk) {
        const int nv = (nvol - v0) < chunk ? (nvol - v0) : chunk;
        const float* cvol = vol + (size_t)v0 * N;
        unsigned* cstats = stats + (size_t)v0 * (X + Y + Z) * 2;
        rc = launch_plane_stats_f32(cvol, nv, X, Y, Z, cstats, stream);
        if (rc) return rc;
        ScatterOuts so;
        memset(&so, 0, sizeof(so));
        for (int pl = 0; pl < 3; ++pl) {
            bool want = false;
            for (int mej = MSL_MEJORA_HE; mej <= MSL_MEJORA_LT; ++mej) want |= outs[(mej - 1) * 3 + pl] != nullptr;
            if (want) { so.u[pl] = U[pl]; so.pitch[pl] = upitch[pl]; }
        }
        rc = launch_norm_scatter(cvol, nv, X, Y, Z, cstats, so, stream);
        if (rc) return rc;
        DensePlane dplanes[3];
        int ndense = 0;
        bool dense_cl = false;
        for (int pl = 0; pl < 3; ++pl) {
            if (!so.u[pl]) continue;
            uint8_t* dst[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
            uintptr_t align = 0;
            for (int mej = MSL_MEJORA_HE; mej <= MSL_MEJORA_LT; ++mej) {
                uint8_t* o = outs[(mej - 1) * 3 + pl];
                dst[mej] = o ? o + (size_t)v0 * N : nullptr;
                align |= reinterpret_cast<uintptr_t>(dst[mej]);
            }
            const int rows = rows_p[pl], cols = cols_p[pl], npx = rows * cols;
            EnhParams p;
            memset(&p, 0, sizeof(p));
            clahe_geometry(rows, cols, p);
            const bool dense_ok = dense_supported(rows, cols, dst[MSL_MEJORA_CLAHE] != nullptr) && (align & 3) == 0;
            if (dense_ok) {
                // the planes of the chunk share one launch (one grid tail instead of three) when they agree on CLAHE
                const bool cl = dst[MSL_MEJORA_CLAHE] != nullptr;
                if (ndense > 0 && cl != dense_cl) {
                    rc = launch_enhance_dense_multi(dplanes, ndense, tables, dtabs, dtabs_bytes, stream);
                    if (rc) return rc;
                    ndense = 0;
                }
                dense_cl = cl;
                DensePlane& q = dplanes[ndense++];
                q.U = U[pl]; q.u_pitch = upitch[pl]; q.nslices = nv * n_p[pl]; q.rows = rows; q.cols = cols;
                q.out_he = dst[MSL_MEJORA_HE]; q.out_clahe = dst[MSL_MEJORA_CLAHE]; q.out_gc = dst[MSL_MEJORA_GC]; q.out_lt = dst[MSL_MEJORA_LT];
                q.th = p.cl_th; q.tw = p.cl_tw; q.clip = p.cl_clip; q.lut_scale = p.cl_lut_scale;
                continue;
            }
            // generic slice kernel on the staged stack (PNG orientation: G[a, b] = P[cols-1-b, a])
            for (int mej = MSL_MEJORA_HE; mej <= MSL_MEJORA_LT; ++mej) {
                if (!dst[mej]) continue;
                p.in = U[pl];
                p.vol_stride = (long long)n_p[pl] * (long long)upitch[pl]; p.idx_stride = (long long)upitch[pl];
                p.base0 = (long long)(cols - 1) * rows;
                p.sa = 1; p.sb = -(long long)rows; p.rows = rows; p.cols = cols; p.nvol = nv; p.n_plane = n_p[pl];
                p.vol_of_slice = nullptr; p.idx_of_slice = nullptr;
                p.out = dst[mej]; p.out_pitch = npx; p.layout = MSL_OUT_P; p.mejora = mej; p.tables = tables;
                rc = launch_enhance_slices(p, MSL_U8, nv * n_p[pl], stream);
                if (rc) return rc;
            }
        }
        if (ndense > 0) {
            rc = launch_enhance_dense_multi(dplanes, ndense, tables, dtabs, dtabs_bytes, stream);
            if (rc) return rc;
        }
    }
    return MSL_OK;
}

int msl_stage_slices(const float* vol, int nvol, int X, int Y, int Z, int plano, const int32_t* vol_of_slice,
                     const int32_t* idx_of_slice, int nslices, uint8_t* out, size_t slice_pitch_bytes, msl_stream_t stream) {
    MSL_REQUIRE(vol && out, "NULL pointer");
    MSL_REQUIRE(nvol > 0 && X > 0 && Y > 0 && Z > 0 && nslices >= 0, "non-positive size");
    MSL_REQUIRE(plano >= MSL_AXIAL && plano <= MSL_SAGITAL, "Plano %d no válido.", plano);
    MSL_REQUIRE((vol_of_slice == nullptr) == (idx_of_slice == nullptr), "vol_of_slice and idx_of_slice must both be given or both be NULL");
    const int n_plane = n_plane_of(plano, X, Y, Z);
    const long long npx = (long long)X * Y * Z / n_plane;
    if (!vol_of_slice)
        MSL_REQUIRE((long long)nslices == (long long)nvol * n_plane, "dense mode needs nslices == nvol * n_plane (%lld), got %d", (long long)nvol * n_plane, nslices);
    MSL_REQUIRE(slice_pitch_bytes >= (size_t)npx, "slice pitch smaller than a slice");
    return launch_stage_slices(vol, nvol, X, Y, Z, plano, vol_of_slice, idx_of_slice, nslices, out, slice_pitch_bytes, (cudaStream_t)stream);
}

size_t msl_enhance_stack_workspace_bytes(int rows, int cols) { return dense_tabs_bytes(rows, cols); }

int msl_enhance_stack(const uint8_t* stack_p, size_t slice_pitch_bytes, int nslices, int rows, int cols,
                      uint8_t* out_he, uint8_t* out_clahe, uint8_t* out_gc, uint8_t* out_lt,
                      const uint8_t* tables, void* ws, size_t ws_bytes, msl_stream_t stream) {
    MSL_REQUIRE(stack_p && tables, "NULL pointer");
    MSL_REQUIRE(nslices >= 0 && rows > 0 && cols > 0, "non-positive size");
    MSL_REQUIRE(out_he || out_clahe || out_gc || out_lt, "no output wanted");
    MSL_REQUIRE(slice_pitch_bytes >= (size_t)rows * cols && (slice_pitch_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(stack_p) & 15) == 0,
                "the staged stack needs a 16-byte aligned base and a slice pitch that is a multiple of 16 bytes");
    if (nslices == 0) return MSL_OK;
    if (!dense_supported(rows, cols, out_clahe != nullptr)) {
        set_error("msl_enhance_stack: %d x %d slices are outside the dense kernel's range (use msl_enhance_images)", rows, cols);
        return MSL_ERR_UNSUPPORTED;
    }
    EnhParams p;
    memset(&p, 0, sizeof(p));
    clahe_geometry(rows, cols, p);
    DensePlane q;
    q.U = stack_p; q.u_pitch = slice_pitch_bytes; q.nslices = nslices; q.rows = rows; q.cols = cols;
    q.out_he = out_he; q.out_clahe = out_clahe; q.out_gc = out_gc; q.out_lt = out_lt;
    q.th = p.cl_th; q.tw = p.cl_tw; q.clip = p.cl_clip; q.lut_scale = p.cl_lut_scale;
    return launch_enhance_dense_multi(&q, 1, tables, ws, ws_bytes, (cudaStream_t)stream);
}

size_t msl_png_bytes(int H, int W, int channels) {
    if (H <= 0 || W <= 0 || (channels != 1 && channels != 4)) return 0;
    return png_file_bytes(H, W, channels);
}

int msl_png_pack(const uint8_t* pixels, int n, int H, int W, int channels, uint8_t* out, size_t out_pitch_bytes, msl_stream_t stream) {
    MSL_REQUIRE(n >= 0 && H > 0 && W > 0, "non-positive size");
    MSL_REQUIRE(channels == 1 || channels == 4, "channels %d no válido (1 o 4)", channels);
    if (n == 0) return MSL_OK;
    MSL_REQUIRE(pixels && out, "NULL pointer");
    MSL_REQUIRE((unsigned long long)H * ((unsigned long long)W * channels + 1) < 0x7fffffffull, "image too large");
    const size_t need = (png_file_bytes(H, W, channels) + 15) & ~(size_t)15;
    MSL_REQUIRE((out_pitch_bytes & 15) == 0 && out_pitch_bytes >= need && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "out must be 16-byte aligned with a pitch that is a multiple of 16 and >= %zu", need);
    return launch_png_pack(pixels, n, H, W, channels, out, out_pitch_bytes, (cudaStream_t)stream);
}

size_t msl_deflate_bound(int n, int container, size_t raw) {
    if (n <= 0 || container < MSL_Z_RAW || container > MSL_Z_PNG) return 0;
    return (size_t)n * deflate_slot_bytes(container, raw);
}

size_t msl_deflate_workspace_bytes(int n, int container, size_t raw) {
    if (n <= 0 || container < MSL_Z_RAW || container > MSL_Z_PNG) return 0;
    return deflate_workspace_bytes(n, container, raw);
}

int msl_deflate_chunks(const uint8_t* src, size_t total_len, size_t chunk_len, int container, int dist2, uint8_t* out, size_t out_cap,
                       uint64_t* out_off, uint32_t* out_meta, void* ws, size_t ws_bytes, msl_stream_t stream) {
    return msl_deflate_files(src, 1, total_len, total_len, nullptr, 0, 0, 0, chunk_len, container, dist2, out, out_cap, out_off, out_meta,
                             ws, ws_bytes, stream);
}

int msl_deflate_files(const uint8_t* bodies, int nfiles, size_t body_pitch, size_t body_len, const uint8_t* prefix, size_t prefix_pitch,
                      size_t prefix_len, int expand_u8_to_f32, size_t chunk_len, int container, int dist2, uint8_t* out, size_t out_cap,
                      uint64_t* out_off, uint32_t* out_meta, void* ws, size_t ws_bytes, msl_stream_t stream) {
    MSL_REQUIRE(container >= MSL_Z_RAW && container <= MSL_Z_GZIP, "container %d no válido (MSL_Z_RAW / ZLIB / GZIP)", container);
    MSL_REQUIRE(chunk_len > 0 && nfiles > 0, "chunk_len and nfiles must be positive");
    MSL_REQUIRE(out && out_off, "NULL out / out_off");
    MSL_REQUIRE(bodies || body_len == 0, "NULL bodies");
    MSL_REQUIRE(prefix || prefix_len == 0, "NULL prefix");
    MSL_REQUIRE(nfiles == 1 || body_pitch >= body_len, "body pitch smaller than a body");
    const size_t total = prefix_len + body_len * (expand_u8_to_f32 ? 4 : 1);
    const size_t spv = total == 0 ? 1 : (total + chunk_len - 1) / chunk_len;
    MSL_REQUIRE(spv * (size_t)nfiles <= 0x7fffffff, "too many chunks");
    return launch_deflate_pack(bodies, (int)(spv * nfiles), body_pitch, chunk_len, total, 0, 0, 0, 0, container, dist2,
                               prefix_len ? prefix : nullptr, prefix_pitch, prefix_len, expand_u8_to_f32 ? 1 : 0, out, out_cap,
                               reinterpret_cast<unsigned long long*>(out_off), out_meta, ws, ws_bytes, (cudaStream_t)stream);
}

int msl_png_encode(const uint8_t* pixels, int n, int H, int W, int channels, uint8_t* out, size_t out_cap, uint64_t* out_off,
                   void* ws, size_t ws_bytes, msl_stream_t stream) {
    MSL_REQUIRE(n >= 0 && H > 0 && W > 0, "non-positive size");
    MSL_REQUIRE(channels >= 1 && channels <= 4, "channels %d no válido (1..4)", channels);
    MSL_REQUIRE(out_off, "NULL out_off");
    if (n == 0) return MSL_OK;
    MSL_REQUIRE(pixels && out, "NULL pointer");
    const size_t rb = (size_t)W * channels;
    MSL_REQUIRE(rb < (1u << 24), "scanline too long");
    return launch_deflate_pack(pixels, n, (size_t)H * rb, 0, 0, H, (int)rb, W, channels, MSL_Z_PNG, channels > 1 ? channels : 0, nullptr, 0, 0, 0,
                               out, out_cap,
                               reinterpret_cast<unsigned long long*>(out_off), nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

int msl_inflate(const uint8_t* src, size_t src_bytes, const uint64_t* src_off, int n, int container, uint8_t* dst, const uint64_t* dst_off,
                uint32_t* status, msl_stream_t stream) {
    MSL_REQUIRE(n >= 0, "negative stream count");
    if (n == 0) return MSL_OK;
    MSL_REQUIRE(src && src_off && dst && dst_off && status, "NULL pointer");
    MSL_REQUIRE(container >= MSL_Z_RAW && container <= MSL_Z_GZIP, "container %d no válido (MSL_Z_RAW / ZLIB / GZIP)", container);
    return launch_inflate(src, src_bytes, reinterpret_cast<const unsigned long long*>(src_off), n, container, dst,
                          reinterpret_cast<const unsigned long long*>(dst_off), status, (cudaStream_t)stream);
}

int msl_png_unfilter(uint8_t* raw, const uint64_t* raw_off, int n, int H, int W, int bytes_per_pixel, uint8_t* out, uint32_t* status,
                     msl_stream_t stream) {
    MSL_REQUIRE(n >= 0 && H > 0 && W > 0, "non-positive size");
    MSL_REQUIRE(bytes_per_pixel >= 1 && bytes_per_pixel <= 4, "bytes_per_pixel %d no válido (1..4)", bytes_per_pixel);
    if (n == 0) return MSL_OK;
    MSL_REQUIRE(raw && raw_off && out && status, "NULL pointer");
    return launch_png_unfilter(raw, reinterpret_cast<const unsigned long long*>(raw_off), n, H, W, bytes_per_pixel, out, status,
                               (cudaStream_t)stream);
}

int msl_nifti_convert(const uint8_t* payload, int datatype, uint64_t nvox, double slope, double inter, int scaled, float* out_f32,
                      uint8_t* out_u8, double* out_f64, uint64_t* inexact, msl_stream_t stream) {
    MSL_REQUIRE(payload && inexact && (out_f32 || out_u8 || out_f64), "NULL pointer");
    switch (datatype) {
        case 2: case 4: case 8: case 16: case 64: case 256: case 512: case 768: break;
        default: set_error("NIfTI datatype %d not supported", datatype); return MSL_ERR_UNSUPPORTED;
    }
    return launch_nifti_convert(payload, datatype, nvox, slope, inter, scaled, out_f32, out_u8, out_f64,
                                reinterpret_cast<unsigned long long*>(inexact), (cudaStream_t)stream);
}

int msl_mask_contours(const uint8_t* masks, int n, int H, int W, int value, int max_contours, int max_points, uint32_t* counts,
                      uint32_t* contour_len, int16_t* points, msl_stream_t stream) {
    MSL_REQUIRE(n >= 0 && H > 0 && W > 0, "non-positive size");
    MSL_REQUIRE(value >= 0 && value <= 255, "value %d outside 0..255", value);
    MSL_REQUIRE(max_contours > 0 && max_points > 0, "non-positive capacity");
    if (n == 0) return MSL_OK;
    MSL_REQUIRE(masks && counts && contour_len && points, "NULL pointer");
    return launch_contours(masks, n, H, W, value, max_contours, max_points, counts, contour_len, points, (cudaStream_t)stream);
}

int msl_nonzero_flags(const uint8_t* stack, int nvol, int A, int B, int C, uint8_t* any_a, uint8_t* any_b, msl_stream_t stream) {
    MSL_REQUIRE(stack && any_a && any_b, "NULL pointer");
    MSL_REQUIRE(nvol > 0 && A > 0 && B > 0 && C > 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    return launch_nonzero_flags(stack, nvol, A, B, C, any_a, any_b, (cudaStream_t)stream);
}

int msl_copy_box_d2h(void* host_dst, const uint8_t* dev_src, int A, int B, int C, int a0, int a1, int b0, int b1, msl_stream_t stream) {
    MSL_REQUIRE(host_dst && dev_src, "NULL pointer");
    MSL_REQUIRE(A > 0 && B > 0 && C > 0, "non-positive size");
    MSL_REQUIRE(0 <= a0 && a0 <= a1 && a1 <= A && 0 <= b0 && b0 <= b1 && b1 <= B, "box [%d, %d) x [%d, %d) outside [0, %d) x [0, %d)", a0, a1, b0, b1, A, B);
    if (a0 == a1 || b0 == b1) return MSL_OK;
    const size_t pitch = (size_t)B * C, off = ((size_t)a0 * B + b0) * C;
    MSL_CUDA_CHECK(cudaMemcpy2DAsync(static_cast<uint8_t*>(host_dst) + off, pitch, dev_src + off, pitch, (size_t)(b1 - b0) * C,
                                     (size_t)(a1 - a0), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return MSL_OK;
}

int msl_copy_boxes_d2h(void* host_dst, const uint8_t* dev_src, int nvol, int A, int B, int C, const int32_t* boxes, msl_stream_t stream) {
    MSL_REQUIRE(boxes && nvol >= 0, "NULL boxes / negative count");
    const size_t vol = (size_t)A * B * C;
    for (int v = 0; v < nvol; ++v) {
        const int rc = msl_copy_box_d2h(static_cast<uint8_t*>(host_dst) + v * vol, dev_src + v * vol, A, B, C,
                                        boxes[4 * v], boxes[4 * v + 1], boxes[4 * v + 2], boxes[4 * v + 3], stream);
        if (rc) return rc;
    }
    return MSL_OK;
}

int msl_bgr_to_gray(const uint8_t* bgr, size_t npx, uint8_t* gray, msl_stream_t stream) {
    if (npx == 0) return MSL_OK;
    MSL_REQUIRE(bgr && gray, "NULL pointer");
    return launch_bgr_to_gray(bgr, npx, gray, (cudaStream_t)stream);
}

int msl_combine_predictions(const float* masks, const int32_t* inst_offset, int nslices, int mh, int mw,
                            int rows, int cols, int layout, uint8_t* out, msl_stream_t stream) {
    MSL_REQUIRE(nslices >= 0 && rows > 0 && cols > 0, "non-positive size");
    MSL_REQUIRE(layout == MSL_OUT_P || layout == MSL_OUT_G, "layout %d no válido (MSL_OUT_P o MSL_OUT_G)", layout);
    if (nslices == 0) return MSL_OK;
    MSL_REQUIRE(inst_offset && out, "NULL inst_offset / out");
    MSL_REQUIRE(nslices <= 65535, "at most 65535 slices per call");
    MSL_REQUIRE(mh > 0 && mw > 0, "non-positive mask size");      // masks may be NULL when no slice has an instance
    return launch_combine_predictions(masks, inst_offset, nslices, mh, mw, rows, cols, layout, out, (cudaStream_t)stream);
}

int msl_recon(const uint8_t* slices, size_t slice_pitch_bytes, const int32_t* vol_of_slice, const int32_t* idx_of_slice,
              int nslices, int plano, int nvol, int X, int Y, int Z, uint8_t* vol_u8, float* vol_f32,
              void* ws, size_t ws_bytes, msl_stream_t stream) {
    MSL_REQUIRE(nvol > 0 && X > 0 && Y > 0 && Z > 0 && nslices >= 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    MSL_REQUIRE(plano >= MSL_AXIAL && plano <= MSL_SAGITAL, "Plano %d no válido.", plano);
    MSL_REQUIRE(vol_u8 || vol_f32, "one of vol_u8 / vol_f32 must be given");
    MSL_REQUIRE(nslices == 0 || (slices && vol_of_slice && idx_of_slice), "NULL slice arrays");
    const int rows = plano == MSL_SAGITAL ? Y : X, cols = plano == MSL_AXIAL ? Y : Z;
    MSL_REQUIRE(nslices == 0 || slice_pitch_bytes >= (size_t)rows * cols, "slice pitch smaller than a %d x %d slice", rows, cols);
    const size_t need = msl_workspace_bytes(MSL_WS_RECON, nvol, X, Y, Z);
    if (!ws || ws_bytes < need) { set_error("workspace of %zu bytes needed, %zu given", need, ws_bytes); return MSL_ERR_WORKSPACE; }
    return launch_recon(slices, slice_pitch_bytes, vol_of_slice, idx_of_slice, nslices, plano, nvol, X, Y, Z,
                        vol_u8, vol_f32, reinterpret_cast<int32_t*>(ws), (cudaStream_t)stream);
}

int msl_consensus_eval(const uint8_t* ax, const uint8_t* co, const uint8_t* sa, const uint8_t* gt, int nvol, size_t nvox,
                       int umbral, uint8_t* consenso, int64_t* counts, msl_stream_t stream) {
    MSL_REQUIRE(ax && co && sa, "NULL plane volume");
    MSL_REQUIRE((gt == nullptr) == (counts == nullptr), "gt and counts must both be given or both be NULL");
    MSL_REQUIRE(gt || consenso, "nothing to compute: neither consenso nor counts requested");
    MSL_REQUIRE(nvol > 0 && nvox > 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    return launch_consensus_eval(ax, co, sa, gt, nvol, nvox, umbral, consenso, reinterpret_cast<long long*>(counts),
                                 (cudaStream_t)stream);
}

int msl_confusion_counts(const uint8_t* gt, const uint8_t* pred, int nvol, size_t nvox, int64_t* counts, msl_stream_t stream) {
    MSL_REQUIRE(gt && pred && counts, "NULL pointer");
    MSL_REQUIRE(nvol > 0 && nvox > 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    return launch_confusion_counts(gt, pred, nvol, nvox, reinterpret_cast<long long*>(counts), (cudaStream_t)stream);
}

int msl_slice_counts(const uint8_t* gt, const uint8_t* pred, int nvol, int X, int Y, int Z, int64_t* counts, msl_stream_t stream) {
    MSL_REQUIRE(gt && pred && counts, "NULL pointer");
    MSL_REQUIRE(nvol > 0 && X > 0 && Y > 0 && Z > 0, "non-positive size");
    MSL_REQUIRE(nvol <= 65535, "at most 65535 volumes per call");
    MSL_REQUIRE((unsigned long long)X * Y * Z < 0xffff0000ull, "volume of %d x %d x %d voxels too large", X, Y, Z);
    MSL_REQUIRE((size_t)(X + Y + Z) * 16 <= 200 * 1024, "too many slices for the shared-memory counters");
    return launch_slice_counts(gt, pred, nvol, X, Y, Z, reinterpret_cast<long long*>(counts), (cudaStream_t)stream);
}

}  // extern "C"
