/*
 * mslesseg.h - C ABI of libmslesseg.so, the B200 (sm_100a) implementation of the voxel-level
 * volume path of srozenblum/YOLO-MSLesSeg.
 *
 * The reference is pure Python and has no FFI of its own: the seam is a set of Python call
 * sites.  Each entry point below names the reference function(s) whose arithmetic it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *  - Every data pointer is a DEVICE pointer owned by the caller (e.g. a torch CUDA tensor).  The
 *    library allocates nothing, frees nothing and keeps no state except a thread-local error string.
 *  - Work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *    default stream).  The caller synchronises.
 *  - Return value: MSL_OK (0) or a negative MSL_ERR_* code; msl_last_error() describes the failure.
 *  - Volumes are C-contiguous [nvol][Z][Y][X] (x fastest) - byte-identical to the Fortran-ordered
 *    (X, Y, Z) array nibabel's get_fdata() yields (reference utils/Paciente.py:168).
 *  - Slice orientation ("G"): the 2-D array the reference sees, S.shape = (rows, cols):
 *       axial   S[x, y] = V[x, y, k]   (rows, cols) = (X, Y)     utils/Paciente.py:240
 *       coronal S[x, z] = V[x, j, z]   (rows, cols) = (X, Z)     utils/Paciente.py:241
 *       sagital S[y, z] = V[i, y, z]   (rows, cols) = (Y, Z)     utils/Paciente.py:242
 *    PNG orientation ("P"): what scripts/extraer_dataset.py:192 saves, imsave(G.T, origin="lower"):
 *       P[r, c] = G[c, cols-1-r],  P.shape = (cols, rows).
 */
#ifndef MSLESSEG_H
#define MSLESSEG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSL_ABI_VERSION 1

#define MSL_OK               0
#define MSL_ERR_ARG         -1   /* NULL pointer, bad enum, non-positive size, bad pitch        */
#define MSL_ERR_UNSUPPORTED -2   /* combination not implemented (see each function)           */
#define MSL_ERR_CUDA        -3   /* a CUDA runtime call failed; message holds cudaGetErrorString */
#define MSL_ERR_WORKSPACE   -4   /* workspace missing or smaller than msl_workspace_bytes()    */

/* planes - reference utils/Paciente.py:68 PLANOS */
#define MSL_AXIAL   0
#define MSL_CORONAL 1
#define MSL_SAGITAL 2

/* enhancements - reference utils/Paciente.py:67 MEJORAS, utils/mejora_imagen.py */
#define MSL_MEJORA_NONE  0
#define MSL_MEJORA_HE    1   /* mejora_imagen.py:52-67   == cv2.equalizeHist on the normalised slice */
#define MSL_MEJORA_CLAHE 2   /* mejora_imagen.py:91-117  == LUT_OUT[clahe_8x8_clip2(LUT_L[u])]        */
#define MSL_MEJORA_GC    3   /* mejora_imagen.py:139-151 == GC_T[u]                                   */
#define MSL_MEJORA_LT    4   /* mejora_imagen.py:166-184 == LT_T[max u][u]                            */

/* element type of an input image / volume */
#define MSL_F32 0
#define MSL_U8  1

/* output layouts of the enhance entry points */
#define MSL_OUT_G        0   /* gray slice, slice orientation (rows, cols), 1 byte / pixel           */
#define MSL_OUT_P        1   /* gray slice, PNG orientation (cols, rows), 1 byte / pixel             */
#define MSL_OUT_PNG_GRAY 2   /* imsave(cmap="gray") gray byte, PNG orientation, 1 byte / pixel (E8)  */
#define MSL_OUT_PNG_RGBA 3   /* imsave RGBA pixels, PNG orientation, 4 bytes / pixel (E8)            */

/* Constant tables, one caller-owned device buffer of MSL_TABLES_BYTES bytes:
 *   [0    ,  256)  LUT_L    gray -> Lab L            (cv2 GRAY2BGR + BGR2LAB,  mejora_imagen.py:98,101)
 *   [256  ,  512)  LUT_OUT  L'   -> gray             (cv2 LAB2BGR + BGR2GRAY,  mejora_imagen.py:112,115; utils.py:426)
 *   [512  ,  768)  GC_T     gamma table              (mejora_imagen.py:146)
 *   [768  , 1024)  CM       matplotlib gray colormap bytes (extraer_dataset.py:192)
 *   [1024 , 1024+65536) LT_T[m][v]  log table for a slice whose maximum is m (mejora_imagen.py:173-182)
 * They are parameters computed on the host with the reference's own NumPy expressions
 * (mslesseg_b200/tables.py); the library only reads them.  LUT_L must be non-decreasing (gray -> L is; the
 * volume path derives CLAHE's L histograms from the gray histograms through it); the other tables are free. */
#define MSL_TAB_LUT_L   0
#define MSL_TAB_LUT_OUT 256
#define MSL_TAB_GC      512
#define MSL_TAB_CM      768
#define MSL_TAB_LT      1024
#define MSL_TABLES_BYTES (1024 + 65536)

/* workspace kinds for msl_workspace_bytes() */
#define MSL_WS_ENHANCE_VOLUMES 1
#define MSL_WS_RECON           2

typedef void* msl_stream_t;   /* cudaStream_t */

int         msl_version(void);
const char* msl_last_error(void);

/* Bytes of scratch the caller must pass as `ws` to the entry point `op` (MSL_WS_*). */
size_t msl_workspace_bytes(int op, int nvol, int X, int Y, int Z);

/* ---- E0: lesion-slice flags -------------------------------------------------------------------
 * Replaces the `np.any(mask_slice > 0)` loop of Paciente.indices_cortes_con_lesion
 * (utils/Paciente.py:252-259) for all three planes in one pass over the mask.
 * gt: [nvol][Z][Y][X], dtype MSL_U8 or MSL_F32.  any_ax[nvol][Z], any_co[nvol][Y], any_sa[nvol][X]
 * receive 1 where the slice holds a voxel > 0, else 0.  The index-window arithmetic of
 * indices_a_usar (:261-275) and the percentile (scripts/extraer_dataset.py:110-135) stay on the host. */
int msl_lesion_slices(const void* gt, int dtype, int nvol, int X, int Y, int Z,
                      uint8_t* any_ax, uint8_t* any_co, uint8_t* any_sa, msl_stream_t stream);

/* ---- E1 statistics: per-slice intensity range of the three planes -------------------------------------
 * The min / ptp that normalizar_a_uint8 (utils/utils.py:400-405) takes per slice, for every slice of every plane in one
 * pass over the volume; also what calcular_rango_global (extras/generar_gif_predicciones.py:141-148: minimum and maximum
 * over a list of slices) reduces.  vol: float32 [nvol][Z][Y][X]; ranges: float32 [nvol][Z + Y + X][2] = {min, max},
 * rows [0, Z) axial, [Z, Z+Y) coronal, [Z+Y, Z+Y+X) sagital.  X <= 256. */
int msl_slice_ranges(const float* vol, int nvol, int X, int Y, int Z, float* ranges, msl_stream_t stream);

/* ---- E1-E8: enhance a list of slices of resident volumes ------------------------------------
 * Replaces, per slice, Paciente.obtener_corte_imagen (utils/Paciente.py:216-222) ->
 * aplicar_mejora (:195-210) -> <HE|CLAHE|GC|LT>.aplicar (utils/mejora_imagen.py) incl.
 * convertir_a_bgr / normalizar_a_uint8 (utils/utils.py:396-418), verificar_grises
 * (utils/utils.py:421-427, called at scripts/extraer_dataset.py:190) and, for the PNG layouts,
 * the orientation + matplotlib normalisation / colormap of guardar_cortes (extraer_dataset.py:192,197).
 *
 * vol: [nvol][Z][Y][X] of `dtype`.  Slice s is plane `plano`, volume vol_of_slice[s], index
 * idx_of_slice[s] (device int32 arrays; both NULL = dense: every index of every volume,
 * s = v * n_plane + i, and nslices must equal nvol * n_plane).  Entries outside the volume are
 * skipped.  Output slice s starts at out + s * slice_pitch_bytes.
 * MSL_F32 input is normalised per slice exactly like normalizar_a_uint8 (utils/utils.py:396-406); MSL_U8
 * input is used as is.  mejora NONE: layouts G / P return that uint8 slice itself (the E1 output; mask
 * slices for U8 input); the PNG layouts on F32 input reproduce imsave of the RAW slice (float64
 * normalisation, what guardar_cortes saves when the experiment has no enhancement). */
int msl_enhance_slices(const void* vol, int dtype, int nvol, int X, int Y, int Z,
                       int mejora, int plano,
                       const int32_t* vol_of_slice, const int32_t* idx_of_slice, int nslices,
                       uint8_t* out, size_t slice_pitch_bytes, int layout,
                       const uint8_t* tables, msl_stream_t stream);

/* ---- E1-E8 on stand-alone 2-D images --------------------------------------------------------
 * Replaces Algoritmo.aplicar(imagen) + verificar_grises for a batch of C-contiguous 2-D images
 * (utils/mejora_imagen.py:31,52,91,139,166).  imgs: [nimg] images of rows x cols elements,
 * image n at imgs + n * img_pitch_elems elements. */
int msl_enhance_images(const void* imgs, int dtype, int nimg, int rows, int cols, size_t img_pitch_elems,
                       int mejora, uint8_t* out, size_t out_pitch_bytes, int layout,
                       const uint8_t* tables, msl_stream_t stream);

/* ---- E1-E7 whole volumes, all three planes from one resident copy ----------------------------
 * Same results as msl_enhance_slices(dense) for every plane, produced by a tri-planar pipeline
 * (per-slice min/max of the three planes in one pass, normalise + scatter, per-slice HE/CLAHE).
 * outs: HOST array of 12 DEVICE pointers indexed (mejora-1)*3 + plano; NULL = not wanted.
 * outs[(m-1)*3+p] receives [nvol][n_p] slices in PNG orientation (MSL_OUT_P), densely packed.
 * ws: msl_workspace_bytes(MSL_WS_ENHANCE_VOLUMES, ...) bytes of device scratch. */
int msl_enhance_volumes(const float* vol, int nvol, int X, int Y, int Z,
                        uint8_t* const* outs, const uint8_t* tables,
                        void* ws, size_t ws_bytes, msl_stream_t stream);

/* ---- E3-E6 on a staged stack of normalised slices ----------------------------------------------
 * The per-slice stage of msl_enhance_volumes on its own: `nslices` uint8 slices that already went through E1 (e.g.
 * msl_enhance_slices with MSL_MEJORA_NONE and layout MSL_OUT_P into a buffer whose slice pitch is a multiple of 16 bytes),
 * PNG orientation [cols][rows], -> any of HE / CLAHE / GC / LT (utils/mejora_imagen.py:52-184), densely packed
 * [nslices][cols][rows], same orientation.  rows x cols = the slice orientation of the plane (what clahe's tile grid is laid
 * over).  Same bytes as msl_enhance_slices produces slice by slice, at the speed of the whole-volume kernel; this is how
 * mslesseg_b200.ops.enhance_slices runs long slice lists (Paciente.cortes_con_lesion_* of a whole cohort).
 * ws: msl_enhance_stack_workspace_bytes(rows, cols) bytes, 16-byte aligned (only read when out_clahe is given).
 * MSL_ERR_UNSUPPORTED when the geometry is outside the dense kernel's range. */
size_t msl_enhance_stack_workspace_bytes(int rows, int cols);
/* E1 + E2 for a slice list, the front half of that route: slice s = (vol_of_slice[s], idx_of_slice[s]) of plane `plano` of
 * the float32 volumes (NULL lists: every slice of every volume) is normalised like normalizar_a_uint8 (utils/utils.py:396-406)
 * and written in PNG orientation to out + s * slice_pitch_bytes.  Pairs outside the volumes are skipped.
 * MSL_ERR_UNSUPPORTED for odd slice rows (use msl_enhance_slices with MSL_MEJORA_NONE). */
int msl_stage_slices(const float* vol, int nvol, int X, int Y, int Z, int plano, const int32_t* vol_of_slice,
                     const int32_t* idx_of_slice, int nslices, uint8_t* out, size_t slice_pitch_bytes, msl_stream_t stream);
int msl_enhance_stack(const uint8_t* stack_p, size_t slice_pitch_bytes, int nslices, int rows, int cols,
                      uint8_t* out_he, uint8_t* out_clahe, uint8_t* out_gc, uint8_t* out_lt,
                      const uint8_t* tables, void* ws, size_t ws_bytes, msl_stream_t stream);

/* ---- E8 container: PNG files around the imsave pixels (encode side) ---------------
 * Replaces the per-slice PNG encode behind plt.imsave (scripts/extraer_dataset.py:192,197).  pixels: uint8
 * [n][H][W][channels], channels 4 (RGBA, colour type 6 - what imsave writes) or 1 (gray, colour type 0).
 * out: n complete PNG files of msl_png_bytes(H, W, channels) bytes each, file i at out + i * out_pitch_bytes
 * (out 16-byte aligned, out_pitch_bytes a multiple of 16 and >= the file size rounded up to 16; the padding is
 * zeroed).  The IDAT chunk holds a zlib stream of STORED deflate blocks: the files decode to exactly `pixels`,
 * compression is left to whoever wants it.  Images whose file exceeds ~200 KB are refused (MSL_ERR_UNSUPPORTED). */
size_t msl_png_bytes(int H, int W, int channels);
int msl_png_pack(const uint8_t* pixels, int n, int H, int W, int channels,
                 uint8_t* out, size_t out_pitch_bytes, msl_stream_t stream);

/* ---- byte-stream codec, encode side: deflate on the device -----------------------
 * Replaces the zlib work behind plt.imsave (scripts/extraer_dataset.py:192,197), cv2.imwrite (utils/utils.py:393,
 * scripts/generar_predicciones.py:153) and nib.save (utils/utils.py:176-177): what crosses PCIe and lands on disk are the
 * compressed files.  A stream is a sequence of deflate blocks (RFC 1951) of two kinds, chosen per 16-byte segment inside
 * tiles of 4 KB: stored blocks for bytes that do not repeat (a fixed Huffman code cannot shrink them) and fixed-Huffman
 * blocks holding run matches at ONE distance `dist2` in 1..4 (0 means 1: repeated bytes; 4 suits RGBA pixels and float32
 * voxels: repeated elements) for segments that lie inside a run; the last block is an empty stored block.  Any inflate
 * implementation reads the result (the tests decode every stream with zlib / gzip / Pillow).  Inside the container
 * asked for:
 *   MSL_Z_RAW   bare deflate            MSL_Z_ZLIB  RFC 1950 (0x78 0x01, Adler-32)
 *   MSL_Z_GZIP  RFC 1952 member; header carries an FEXTRA subfield 'M','S' = {u32 member bytes, u32 raw bytes}, so a
 *               file that is a sequence of such members can be split without decoding (msl_inflate runs them in parallel)
 *   MSL_Z_PNG   a complete PNG file (signature, IHDR, one IDAT, IEND), filter type 0 on every scanline
 * The n streams are written back to back: stream i occupies out[out_off[i], out_off[i+1]) (out_off: DEVICE uint64
 * [n + 1]); out_off[n] > out_cap means the capacity was too small (nothing past it was written).
 * msl_deflate_bound() is a capacity that always suffices.  out_meta (optional, DEVICE uint32 [n][4]) receives per stream
 * {container bytes, raw bytes, Adler-32 / CRC-32 of the raw bytes, 0}.  ws: msl_deflate_workspace_bytes(...), 16-byte aligned.
 *
 * msl_deflate_chunks: `total_len` bytes at src cut into n = ceil(total_len / chunk_len) streams (chunk_len < 16 MB).
 * msl_deflate_files : nfiles files at once, file f = prefix f (prefix_len bytes at prefix + f * prefix_pitch; pitch 0 = one
 *                     prefix for all: a NIfTI header) followed by body f (body_len bytes at bodies + f * body_pitch), each
 *                     cut into ceil(file bytes / chunk_len) streams; streams are numbered file-major.  With
 *                     expand_u8_to_f32 the body is a uint8 mask and the file holds it as float32 0.0f / 1.0f (what
 *                     reconstruir_volumen saves, scripts/reconstruir_volumen.py:202) without that volume ever existing.
 * msl_png_encode    : pixels uint8 [n][H][W][channels] (channels 1, 2, 3 or 4) -> n PNG files. */
#define MSL_Z_RAW  0
#define MSL_Z_ZLIB 1
#define MSL_Z_GZIP 2
#define MSL_Z_PNG  3
size_t msl_deflate_bound(int n, int container, size_t raw_len_per_stream);
size_t msl_deflate_workspace_bytes(int n, int container, size_t raw_len_per_stream);
int msl_deflate_chunks(const uint8_t* src, size_t total_len, size_t chunk_len, int container, int dist2,
                       uint8_t* out, size_t out_cap, uint64_t* out_off, uint32_t* out_meta,
                       void* ws, size_t ws_bytes, msl_stream_t stream);
int msl_deflate_files(const uint8_t* bodies, int nfiles, size_t body_pitch, size_t body_len,
                      const uint8_t* prefix, size_t prefix_pitch, size_t prefix_len, int expand_u8_to_f32,
                      size_t chunk_len, int container, int dist2,
                      uint8_t* out, size_t out_cap, uint64_t* out_off, uint32_t* out_meta,
                      void* ws, size_t ws_bytes, msl_stream_t stream);
int msl_png_encode(const uint8_t* pixels, int n, int H, int W, int channels,
                   uint8_t* out, size_t out_cap, uint64_t* out_off,
                   void* ws, size_t ws_bytes, msl_stream_t stream);

/* ---- byte-stream codec, decode side: inflate on the device -----------------------
 * Replaces the zlib work behind nib.load(...).get_fdata() (utils/Paciente.py:168,179, utils/utils.py:156), Image.open
 * (scripts/reconstruir_volumen.py:141) and cv2.imread (utils/utils.py:391): the host uploads FILE bytes.
 * msl_inflate: n independent streams, ONE WARP each (stored, fixed and dynamic Huffman blocks; any conforming deflate
 * stream).  Stream i = src[src_off[i], src_off[i+1]) in container MSL_Z_RAW / MSL_Z_ZLIB / MSL_Z_GZIP (a gzip stream may
 * hold several members, decoded one after the other); its output goes to dst[dst_off[i], dst_off[i+1]) (the capacity).
 * src must be 4-byte aligned; src_off / dst_off are DEVICE uint64 [n + 1].  status: DEVICE uint32 [n][4] =
 * {0 or an error code (1 header, 2 block, 3 code, 4 distance, 5 out of space, 6 input ended), bytes produced, the
 * Adler-32 / CRC-32 stored in the (last) trailer, the stored ISIZE (gzip)}.  Parallelism = number of streams: a file
 * written by msl_deflate_chunks(MSL_Z_GZIP) splits into its members (the 'MS' subfield gives their sizes).
 *
 * msl_png_unfilter: undoes the PNG scanline filters (types 0-4) of n inflated images of H x W pixels, bytes_per_pixel
 * in {1, 2, 3, 4} (8-bit gray, gray+alpha, RGB, RGBA; not interlaced), image i at raw[raw_off[i], raw_off[i+1]), in place,
 * and writes the FIRST channel to out [n][H][W] - what cargar_y_preprocesar_imagen keeps (scripts/reconstruir_volumen.py:
 * 141-145) and msl_recon consumes.  status: DEVICE uint32 [n], 0 = ok.
 *
 * msl_nifti_convert: nvox voxels of NIfTI datatype code `datatype` (2 uint8, 4 int16, 8 int32, 16 float32, 64 float64,
 * 256 int8, 512 uint16, 768 uint32; little endian, any alignment), scaled by slope / inter when scaled != 0 (get_fdata
 * semantics, float64 arithmetic), to out_f32 / out_u8 / out_f64 (any may be NULL).  *inexact (DEVICE uint64, zeroed by the caller) is
 * incremented by the number of values the float32 / uint8 output could not hold exactly (float64 holds every value get_fdata() yields). */
int msl_inflate(const uint8_t* src, size_t src_bytes, const uint64_t* src_off, int n, int container,
                uint8_t* dst, const uint64_t* dst_off, uint32_t* status, msl_stream_t stream);
int msl_png_unfilter(uint8_t* raw, const uint64_t* raw_off, int n, int H, int W, int bytes_per_pixel,
                     uint8_t* out, uint32_t* status, msl_stream_t stream);
int msl_nifti_convert(const uint8_t* payload, int datatype, uint64_t nvox, double slope, double inter, int scaled,
                      float* out_f32, uint8_t* out_u8, double* out_f64, uint64_t* inexact, msl_stream_t stream);

/* ---- E9: external contours of binary masks (YOLO label polygons) ----------------------------
 * Replaces the contour extraction behind anotar_mascaras (scripts/extraer_dataset.py:215-227): ultralytics'
 * convert_segment_masks_to_yolo_seg runs cv2.findContours((mask == value), RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) on every
 * mask PNG and writes one "class x1 y1 x2 y2 ..." line per contour with at least 3 points.
 * masks: uint8 [n][H][W].  Foreground = (byte == value), or any non-zero byte when value == 0.  Per mask the contours come
 * out in OpenCV's order with OpenCV's points (x, y as int16), back to back in points[i][...]:
 *   counts      uint32 [n][4]  = {contours stored, points stored, overflow flags (1: > max_contours, 2: > max_points),
 *                                 contours found}
 *   contour_len uint32 [n][max_contours]  points of each contour
 *   points      int16  [n][max_points][2]
 * The normalisation (x / width, y / height, 6 decimals) and the text stay on the host.  Masks up to ~45,000 pixels
 * ((H + 2) * (W + 2) * 5 bytes of shared memory). */
int msl_mask_contours(const uint8_t* masks, int n, int H, int W, int value, int max_contours, int max_points,
                      uint32_t* counts, uint32_t* contour_len, int16_t* points, msl_stream_t stream);

/* ---- host hand-off helpers: copy only the non-zero box of a result ------------------------------
 * Skull-stripped volumes are two thirds background and predicted masks ~99 % zeros, and the device-to-host copy of the
 * results is what bounds the path end to end.  msl_nonzero_flags marks which slices (a) and rows (b) of a uint8 stack
 * [nvol][A][B][C] hold a non-zero byte (any_a [nvol][A], any_b [nvol][B], overwritten); msl_copy_box_d2h copies the box
 * a0 <= a < a1, b0 <= b < b1 (full rows of C bytes) of ONE [A][B][C] array from the device to a host array of the same
 * shape (cudaMemcpy2DAsync); the caller keeps the rest of the host array zero. */
int msl_nonzero_flags(const uint8_t* stack, int nvol, int A, int B, int C,
                      uint8_t* any_a, uint8_t* any_b, msl_stream_t stream);
int msl_copy_box_d2h(void* host_dst, const uint8_t* dev_src, int A, int B, int C,
                     int a0, int a1, int b0, int b1, msl_stream_t stream);
/* the same for nvol consecutive arrays; boxes: HOST array [nvol][4] = {a0, a1, b0, b1} */
int msl_copy_boxes_d2h(void* host_dst, const uint8_t* dev_src, int nvol, int A, int B, int C,
                       const int32_t* boxes, msl_stream_t stream);

/* ---- E7 helper: verificar_grises (utils/utils.py:421-427) on 3-channel images ------------------
 * cv2.cvtColor(BGR2GRAY) in OpenCV's 8-bit fixed point: (3735 B + 19235 G + 9798 R + 2^14) >> 15.
 * bgr: uint8 [npx][3] interleaved, gray: uint8 [npx]. */
int msl_bgr_to_gray(const uint8_t* bgr, size_t npx, uint8_t* gray, msl_stream_t stream);

/* ---- R0: YOLO instance masks -> predicted slice masks (the producer of msl_recon's input) -------
 * Replaces combinar_predicciones (scripts/generar_predicciones.py:123-133: every instance mask > 0.5,
 * cv2.resize INTER_NEAREST to the image shape, maximum over the instances) and normalizar_prediccion
 * (:136-140: cv2.flip(pred.T, 1) * 255).
 * masks: float32 [n_inst][mh][mw] (ultralytics `masks.data` of all slices, concatenated);
 * inst_offset: int32 [nslices + 1] (device): the instances of slice s are [inst_offset[s], inst_offset[s+1]).
 * The model saw the PNG-oriented image, (height, width) = (cols, rows).  out: uint8, dense, one image per slice:
 *   layout MSL_OUT_P : (cols, rows), values {0, 1}    == combinar_predicciones(preds, (cols, rows))
 *   layout MSL_OUT_G : (rows, cols), values {0, 255}  == normalizar_prediccion(combinar_predicciones(...)),
 *                      i.e. exactly the `slices` argument of msl_recon. */
int msl_combine_predictions(const float* masks, const int32_t* inst_offset, int nslices, int mh, int mw,
                            int rows, int cols, int layout, uint8_t* out, msl_stream_t stream);

/* ---- R1-R2: stack predicted 2-D masks back into volumes ---------------------------------------
 * Replaces cargar_y_preprocesar_imagen's binarisation (scripts/reconstruir_volumen.py:146-148),
 * insertar_corte (:179-186) and the zero-initialised volume of reconstruir_volumen (:199-213).
 * slices: uint8 pred masks in slice orientation (rows, cols), slice s at slices + s*slice_pitch_bytes;
 * voxel = (pixel > 0).  Indices never listed stay 0; validar_corte's range/shape checks (:153-176)
 * are done by the host wrapper, out-of-range entries are skipped here.  If an index is listed
 * twice the slice with the larger s wins.  At least one of vol_u8 / vol_f32 ([nvol][Z][Y][X]) must
 * be non-NULL.  ws: msl_workspace_bytes(MSL_WS_RECON, ...). */
int msl_recon(const uint8_t* slices, size_t slice_pitch_bytes,
              const int32_t* vol_of_slice, const int32_t* idx_of_slice, int nslices, int plano,
              int nvol, int X, int Y, int Z, uint8_t* vol_u8, float* vol_f32,
              void* ws, size_t ws_bytes, msl_stream_t stream);

/* ---- R3-R4: tri-planar vote fused with the voxel confusion counts -----------------------------
 * Replaces combinar_volumenes (scripts/generar_consenso.py:106-109): consenso = (ax+co+sa >= umbral),
 * and the boolean sums behind DSC / precision / recall / AUC (utils/utils.py:455-495) for the three
 * planes and the consensus in ONE pass.  All volumes uint8 [nvol][nvox].
 * counts: int64 [nvol][4][4] = {axial, coronal, sagital, consenso} x {tp, fp, fn, tn} with the
 * reference's exact predicates (gt==1 & p==1, gt==0 & p==1, gt==1 & p==0, gt==0 & p==0); the
 * buffer is overwritten.  tp+fp+fn+tn < nvox reveals non-binary input to the host.
 * gt and counts may both be NULL (vote only); consenso may be NULL (counts only). */
int msl_consensus_eval(const uint8_t* ax, const uint8_t* co, const uint8_t* sa, const uint8_t* gt,
                       int nvol, size_t nvox, int umbral,
                       uint8_t* consenso, int64_t* counts, msl_stream_t stream);

/* ---- R4 for one prediction volume per patient -----------------------------------------------
 * Replaces the sums of generar_diccionario_metricas(gt_vol, pred_vol) (scripts/eval.py:115-128).
 * counts: int64 [nvol][4] = {tp, fp, fn, tn}, overwritten. */
int msl_confusion_counts(const uint8_t* gt, const uint8_t* pred, int nvol, size_t nvox,
                         int64_t* counts, msl_stream_t stream);

/* ---- R4 at slice granularity ----------------------------------------------------
 * The counts behind the per-slice DSC of extras/visualizar_prediccion_corte.py:150-182 (seleccionar_mejor_corte:
 * DSC(pred_slice, gt_slice) for every slice, best one wins), for ALL slices of the three planes in one pass.
 * gt, pred: uint8 [nvol][Z][Y][X].  counts: int64 [nvol][Z + Y + X][4] = {tp, fp, fn, tn} with the same exact
 * ==1 / ==0 predicates as msl_confusion_counts; rows [0, Z) axial, [Z, Z+Y) coronal, [Z+Y, Z+Y+X) sagital. */
int msl_slice_counts(const uint8_t* gt, const uint8_t* pred, int nvol, int X, int Y, int Z,
                     int64_t* counts, msl_stream_t stream);

/* ---- instrumentation (not part of the reference surface) ---------------------------------------
 * msl_kernel_launches: kernels launched by this library since load (total; per kind if non-NULL,
 * array of msl_kernel_kinds() entries).  msl_profile_enable(1) makes every launch record a pair of
 * CUDA events on its stream; msl_profile_collect synchronises them, returns milliseconds and launch
 * counts per kind and switches profiling off.  bench.py uses it for the roofline line.
 * msl_profile_timeline (before the collect): up to `cap` recorded launches in launch order - kind, stream
 * handle, start and end in milliseconds after the first recorded launch; returns the count (-1: error). */
int                msl_kernel_kinds(void);
const char*        msl_kernel_name(int kind);
unsigned long long msl_kernel_launches(unsigned long long* per_kind);
int                msl_profile_enable(int on);
/* Self-test of the E1 division (utils/utils.py:403, f32 (corte - min) / ptp): the kernels replace it by a reciprocal hoisted
 * per slice + a correction step; this entry runs that sequence on n device (g, p) pairs with 1e-18 <= p <= 1e18 and
 * 0 <= g <= p (others are skipped) and returns out4 = {normal quotients that differ from the IEEE division, output bytes that
 * differ, bytes of the packed two-voxel path that differ from the scalar one, pairs looked at} (device, uint64[4]). */
int                msl_selftest_norm_division(const float* g, const float* p, size_t n, unsigned long long* out4, msl_stream_t stream);
int                msl_profile_timeline(int cap, int* kind, unsigned long long* stream, double* start_ms, double* end_ms);
int                msl_profile_collect(double* ms_per_kind, unsigned long long* n_per_kind);

#ifdef __cplusplus
}
#endif
#endif /* MSLESSEG_H */
