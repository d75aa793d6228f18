// Internal launcher interface between the C ABI (msl_abi.cu) and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace msl {

// Kernel kinds for launch accounting / opt-in event timing (msl_profile.cu).
enum KernelKind {
    K_ENH_F32 = 0,        // + mejora (0..4)
    K_ENH_U8 = 5,         // + mejora (0..4)
    K_INIT_STATS = 10,
    K_PLANE_STATS = 11,
    K_LESION_FLAGS = 12,
    K_NORM_SCATTER = 13,
    K_RECON_FILL = 14,
    K_RECON_SLOT_MAP = 15,
    K_RECON_GATHER = 16,
    K_CONSENSUS_EVAL = 17,
    K_CONFUSION_COUNTS = 18,
    K_ENH_DENSE = 19,     // fused HE + CLAHE + GC + LT over staged uint8 stacks
    K_COMBINE_PRED = 20,  // YOLO instance masks -> predicted slice mask
    K_SLICE_COUNTS = 21,  // per-slice confusion counts of the three planes
    K_BGR2GRAY = 22,
    K_PNG_PACK = 23,
    K_NONZERO_FLAGS = 24,
    K_ENH_DENSE_TABLES = 25,   // per-launch plane tables of enhance_dense
    K_DEFLATE = 26,            // fixed-Huffman deflate of byte streams (PNG / zlib / gzip containers)
    K_DEFLATE_SCAN = 27,
    K_DEFLATE_PACK = 28,
    K_INFLATE = 29,
    K_PNG_UNFILTER = 30,
    K_NIFTI_CONVERT = 31,
    K_CHECKSUM = 32,
    K_CONTOURS = 33,
    K_STAGE_SLICES = 34,
    K_NKIND = 35
};

// RAII: counts the launch and, when profiling is enabled, brackets it with CUDA events on `stream`.
class ProfScope {
public:
    ProfScope(int kind, cudaStream_t stream);
    ~ProfScope();
    ProfScope(const ProfScope&) = delete;
    ProfScope& operator=(const ProfScope&) = delete;
private:
    int kind_;
    cudaStream_t stream_;
    cudaEvent_t e0_, e1_;
};

// One launch of the slice-mode enhancement kernel: `nslices` CTAs, one slice each.
struct EnhParams {
    const void* in;              // element type given by the launcher's dtype
    long long vol_stride;        // elements between volumes / images
    long long idx_stride;        // elements between consecutive slice indices of the plane
    long long base0;             // extra element offset (negative-stride views)
    long long sa, sb;            // element strides of slice coordinates (a = row, b = col)
    int rows, cols;              // slice shape in slice orientation
    int nvol, n_plane;           // bounds for (vol_of_slice, idx_of_slice)
    const int32_t* vol_of_slice; // device, or NULL for dense (s = v * n_plane + i)
    const int32_t* idx_of_slice;
    uint8_t* out;
    size_t out_pitch;            // bytes between output slices
    int layout;                  // MSL_OUT_*
    int mejora;                  // MSL_MEJORA_*
    const uint8_t* tables;       // MSL_TABLES_BYTES device bytes
    // CLAHE geometry (host-computed exactly like OpenCV's CLAHE_Impl::apply)
    int cl_th, cl_tw, cl_clip;
    float cl_lut_scale;
    int hist_bytes;              // filled by the launcher
};

size_t enhance_slices_smem_bytes(int mejora, int npx);
int launch_enhance_slices(EnhParams p, int dtype, int nslices, cudaStream_t stream);

// Tri-planar per-slice min/max of float volumes (keys are order-preserving uint32, see f2key).
// stats: [nvol][Z + Y + X][2] = {min key, max key}; must be pre-initialised by init_stats.
int launch_init_stats(unsigned* stats, size_t nslices_total, cudaStream_t stream);
int launch_plane_stats_f32(const float* vol, int nvol, int X, int Y, int Z, unsigned* stats, cudaStream_t stream);
int launch_stage_slices(const float* vol, int nvol, int X, int Y, int Z, int plano, const int32_t* vol_of_slice, const int32_t* idx_of_slice,
                        int nslices, uint8_t* out, size_t out_pitch, cudaStream_t stream);
int launch_selftest_norm_division(const float* g, const float* p, size_t n, unsigned long long* out, cudaStream_t stream);
int launch_stats_keys_to_float(unsigned* stats, size_t n, cudaStream_t stream);    // order-preserving keys -> float bits, in place

// E0: any(voxel > 0) per slice for the three planes.
int launch_lesion_flags(const void* gt, int dtype, int nvol, int X, int Y, int Z,
                        uint8_t* any_ax, uint8_t* any_co, uint8_t* any_sa, cudaStream_t stream);

// Normalise each voxel with the (min, ptp) of the three slices it belongs to and scatter the bytes into
// three PNG-oriented uint8 slice stacks (axial, coronal, sagital); NULL = skip that plane.
struct ScatterOuts {
    uint8_t* u[3];
    size_t pitch[3];             // bytes between consecutive slices of that stack
};
int launch_norm_scatter(const float* vol, int nvol, int X, int Y, int Z, const unsigned* stats,
                        const ScatterOuts& outs, cudaStream_t stream);

// Dense HE / CLAHE over PNG-oriented uint8 stacks (msl_enhance_dense.cu)
size_t dense_u_pitch(int npx);
size_t dense_smem_bytes(int rows, int cols, bool clahe);
bool dense_supported(int rows, int cols, bool clahe);
struct DensePlane {
    const uint8_t* U; size_t u_pitch; int nslices, rows, cols;
    uint8_t *out_he, *out_clahe, *out_gc, *out_lt;    // each may be NULL
    int th, tw, clip; float lut_scale;                // CLAHE geometry of a rows x cols slice
};
// up to three stacks in ONE launch (they must agree on whether CLAHE is wanted)
// tabs_ws: device scratch for the per-plane tables, sum of dense_tabs_bytes(rows, cols) over the CLAHE stacks, 16-byte aligned
size_t dense_tabs_bytes(int rows, int cols);
int launch_enhance_dense_multi(const DensePlane* planes, int nplanes, const uint8_t* tables, void* tabs_ws, size_t tabs_ws_bytes,
                               cudaStream_t stream);

// R1-R2
int launch_recon(const uint8_t* slices, size_t slice_pitch, const int32_t* vol_of_slice, const int32_t* idx_of_slice,
                 int nslices, int plano, int nvol, int X, int Y, int Z, uint8_t* vol_u8, float* vol_f32,
                 int32_t* slot_of, cudaStream_t stream);

// R3-R4
int launch_combine_predictions(const float* masks, const int32_t* inst_offset, int nslices, int mh, int mw, int rows, int cols,
                               int layout, uint8_t* out, cudaStream_t stream);
int launch_bgr_to_gray(const uint8_t* bgr, size_t npx, uint8_t* gray, cudaStream_t stream);
size_t png_file_bytes(int H, int W, int ch);
int launch_png_pack(const uint8_t* pixels, int n, int H, int W, int ch, uint8_t* out, size_t out_pitch, cudaStream_t stream);
// msl_codec.cu: deflate streams in zlib / gzip / PNG containers, packed back to back
size_t deflate_slot_bytes(int container, size_t raw);
size_t deflate_workspace_bytes(int n, int container, size_t raw);
// the gzip member deflate_kernel writes for a chunk of n zero bytes at match distance d (n = 1..4 whole 4 KB tiles): meta =
// {member bytes, n, CRC-32, 0}, bytes into out[kZeroTmplBytes] (host side, cached)
constexpr int kZeroTmplBytes = 400;
void zero_chunk_gzip(int d, unsigned n, uint32_t meta[4], uint8_t* out);
int launch_deflate_pack(const uint8_t* src, int n, size_t src_pitch, size_t chunk, size_t total, int rows, int row_bytes,
                        int img_w, int img_ch, int container, int dist2, const uint8_t* prefix, size_t prefix_pitch, size_t prefix_len,
                        int expand, uint8_t* out, size_t out_cap, unsigned long long* out_off,
                        uint32_t* out_meta, void* ws, size_t ws_bytes, cudaStream_t stream);
// msl_inflate.cu: inflate (one warp per stream), PNG unfilter, NIfTI payload conversion
int launch_inflate(const uint8_t* src, size_t src_bytes, const unsigned long long* src_off, int n, int container, uint8_t* dst,
                   const unsigned long long* dst_off, uint32_t* status, cudaStream_t stream);
int launch_png_unfilter(uint8_t* raw, const unsigned long long* raw_off, int n, int H, int W, int bpp, uint8_t* out, uint32_t* status,
                        cudaStream_t stream);
int launch_nifti_convert(const uint8_t* payload, int datatype, unsigned long long nvox, double slope, double inter, int scaled,
                         float* out_f32, uint8_t* out_u8, double* out_f64, unsigned long long* inexact, cudaStream_t stream);
// msl_contours.cu: external contours of binary masks (cv2.findContours RETR_EXTERNAL / CHAIN_APPROX_SIMPLE)
size_t contours_smem_bytes(int H, int W);
int launch_contours(const uint8_t* masks, int n, int H, int W, int value, int max_contours, int max_points, uint32_t* counts,
                    uint32_t* contour_len, short* points, cudaStream_t stream);
int launch_nonzero_flags(const uint8_t* stack, int nvol, int A, int B, int C, uint8_t* any_a, uint8_t* any_b, cudaStream_t stream);
int launch_slice_counts(const uint8_t* gt, const uint8_t* pred, int nvol, int X, int Y, int Z, long long* counts, cudaStream_t stream);
int launch_consensus_eval(const uint8_t* ax, const uint8_t* co, const uint8_t* sa, const uint8_t* gt,
                          int nvol, size_t nvox, int umbral, uint8_t* consenso, long long* counts, cudaStream_t stream);
int launch_confusion_counts(const uint8_t* gt, const uint8_t* pred, int nvol, size_t nvox, long long* counts,
                            cudaStream_t stream);

}  // namespace msl
