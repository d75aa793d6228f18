// PNG container around the imsave pixels (encode side): replaces the per-slice PNG encode of
// scripts/extraer_dataset.py:192,197 (plt.imsave -> Pillow -> zlib) by one CTA per image that assembles a complete PNG
// file - signature, IHDR, one IDAT whose zlib stream consists of STORED deflate blocks, IEND - in shared memory,
// computes Adler-32 over the scanlines and CRC-32 over the IDAT chunk in parallel, and streams the file out.
// The files decode to exactly the input pixels; compression is left out on purpose (byte work at HBM speed, 4 x the
// size of a deflated file for these images).  Formats: PNG 1.2, RFC 1950 (zlib), RFC 1951 (stored blocks).
#include <cstring>

#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {

constexpr int kPngThreads = 1024;
constexpr uint32_t kCrcPoly = 0xedb88320u;      // reflected CRC-32 polynomial
constexpr int kZOff = 41;                       // 8 signature + 25 IHDR chunk + 4 IDAT length + 4 "IDAT"
constexpr unsigned kStoredMax = 65535u;

// a * b mod P over GF(2), operands in the reflected representation CRC-32 uses (bit 31 = x^0)
__host__ __device__ inline uint32_t multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ kCrcPoly : b >> 1;
        if (m == 0) break;
    }
    return p;
}
// x^(8 n) mod P
inline uint32_t xpow8(unsigned long long n) {
    uint32_t r = 1u << 31, b = 0x00800000u;     // x^0, x^8
    while (n) { if (n & 1) r = multmodp(r, b); b = multmodp(b, b); n >>= 1; }
    return r;
}
inline uint32_t crc32_host(const uint8_t* d, size_t n) {
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; ++i) { c ^= d[i]; for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1; }
    return ~c;
}

struct PngArgs {
    const uint8_t* pixels;
    uint8_t* out;
    size_t out_pitch;
    int H, W, ch;
    unsigned raw;            // H * (1 + W * ch) scanline bytes
    unsigned zlen;           // zlib stream length
    unsigned fsize;          // file size
    unsigned K;              // CRC bytes per thread
    uint32_t M[10];          // x^(8 K 2^l), l = 0 .. 9
    uint32_t crc_init_term;  // multmodp(x^(8 clen), 0xffffffff): what the all-ones initial register turns into
    uint8_t head[kZOff];     // signature, IHDR chunk, IDAT length and type
};

__global__ void __launch_bounds__(kPngThreads) png_pack_kernel(const PngArgs a) {
    extern __shared__ __align__(16) uint8_t F[];              // the file, then the CRC table and the partial CRCs
    const unsigned fpad = (a.fsize + 15u) & ~15u;
    uint32_t* table = reinterpret_cast<uint32_t*>(F + fpad);
    uint32_t* part = table + 256;
    __shared__ unsigned long long s_a[kPngThreads / 32], s_b[kPngThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* px = a.pixels + (size_t)blockIdx.x * a.H * a.W * a.ch;
    if (tid < 256) {
        uint32_t c = (uint32_t)tid;
#pragma unroll
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        table[tid] = c;
    }
    if (tid < kZOff) F[tid] = a.head[tid];
    if (tid == 0) { F[kZOff] = 0x78; F[kZOff + 1] = 0x01; }
    const unsigned nblk = a.raw ? (a.raw + kStoredMax - 1) / kStoredMax : 1;
    if (tid < (int)nblk) {                                    // stored-block headers: BFINAL | BTYPE = 00, LEN, NLEN
        const unsigned len = min(kStoredMax, a.raw - tid * kStoredMax);
        uint8_t* h = F + kZOff + 2 + (size_t)tid * (kStoredMax + 5);
        h[0] = (tid == (int)nblk - 1) ? 1 : 0;
        h[1] = (uint8_t)len; h[2] = (uint8_t)(len >> 8); h[3] = (uint8_t)~len; h[4] = (uint8_t)((~len) >> 8);
    }
    // scanlines: filter byte 0, then the pixel bytes; Adler-32 partial sums on the way.  A warp per scanline: no division
    // per byte, and a scanline (< 65535 bytes, checked by the launcher) touches at most two stored blocks.
    const unsigned rowlen = 1u + (unsigned)a.W * a.ch;
    unsigned long long sa = 0, sb = 0;
    const bool words = a.ch == 4 && (reinterpret_cast<uintptr_t>(px) & 3) == 0;
    for (unsigned y = warp; y < (unsigned)a.H; y += kPngThreads / 32) {
        const unsigned base = y * rowlen, k0 = base / kStoredMax, next = (k0 + 1) * kStoredMax;
        uint8_t* dst = F + kZOff + 2 + 5 * (k0 + 1);          // + idx, + 5 more behind the block boundary
        auto put = [&](unsigned idx, unsigned v) {
            dst[idx + (idx >= next ? 5u : 0u)] = (uint8_t)v;
            sa += v;
            sb += (unsigned long long)(a.raw - idx) * v;
        };
        if (lane == 0) put(base, 0);
        const uint8_t* row = px + (size_t)y * (rowlen - 1);
        if (words) {
            const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row);
            for (unsigned x = lane; x < (unsigned)a.W; x += 32) {
                const uint32_t w = __ldg(row4 + x);
                const unsigned idx = base + 1 + 4 * x;
                put(idx, w & 0xffu); put(idx + 1, (w >> 8) & 0xffu); put(idx + 2, (w >> 16) & 0xffu); put(idx + 3, w >> 24);
            }
        } else {
            for (unsigned xb = lane; xb + 1 < rowlen; xb += 32) put(base + 1 + xb, __ldg(row + xb));
        }
    }
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(FULL, sa, o); sb += __shfl_xor_sync(FULL, sb, o); }
    if (lane == 0) { s_a[warp] = sa; s_b[warp] = sb; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long ta = 1, tb = a.raw;                // a starts at 1, and that 1 is counted once per byte in b
        for (int w = 0; w < kPngThreads / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
        const uint32_t adler = (uint32_t)((tb % 65521ull) << 16) | (uint32_t)(ta % 65521ull);
        uint8_t* q = F + kZOff + a.zlen - 4;
        q[0] = (uint8_t)(adler >> 24); q[1] = (uint8_t)(adler >> 16); q[2] = (uint8_t)(adler >> 8); q[3] = (uint8_t)adler;
    }
    __syncthreads();
    // CRC-32 of the IDAT chunk (type + data): every thread takes K bytes, the last thread the last K; partial registers
    // start from zero, so the missing bytes in front of the first thread are harmless, and partials combine linearly
    const unsigned clen = 4 + a.zlen;
    const uint8_t* C = F + kZOff - 4;
    {
        const long long start = (long long)clen - (long long)(kPngThreads - tid) * a.K;
        uint32_t c = 0;
        for (long long i = start < 0 ? 0 : start; i < start + (long long)a.K; ++i) c = table[(c ^ C[i]) & 0xffu] ^ (c >> 8);
        part[tid] = c;
    }
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < 10; ++l) {
        const int s = 1 << l;
        if (tid < (kPngThreads >> (l + 1))) {
            const int left = (2 * tid + 1) * s - 1, right = (2 * tid + 2) * s - 1;
            part[right] = multmodp(a.M[l], part[left]) ^ part[right];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t crc = part[kPngThreads - 1] ^ a.crc_init_term ^ 0xffffffffu;
        uint8_t* q = F + kZOff + a.zlen;
        q[0] = (uint8_t)(crc >> 24); q[1] = (uint8_t)(crc >> 16); q[2] = (uint8_t)(crc >> 8); q[3] = (uint8_t)crc;
        const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
        for (int i = 0; i < 12; ++i) q[4 + i] = iend[i];
        for (unsigned i = a.fsize; i < fpad; ++i) F[i] = 0;
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.out + (size_t)blockIdx.x * a.out_pitch);
    const uint4* src = reinterpret_cast<const uint4*>(F);
    for (unsigned q = tid; q < fpad / 16; q += kPngThreads) dst[q] = src[q];
}

inline unsigned png_zlen(unsigned raw) { return 2 + 5 * (raw ? (raw + kStoredMax - 1) / kStoredMax : 1) + raw + 4; }

}  // namespace

size_t png_file_bytes(int H, int W, int ch) {
    const unsigned raw = (unsigned)H * (1u + (unsigned)W * ch);
    return (size_t)kZOff + png_zlen(raw) + 4 + 12;
}

int launch_png_pack(const uint8_t* pixels, int n, int H, int W, int ch, uint8_t* out, size_t out_pitch, cudaStream_t stream) {
    PngArgs a;
    memset(&a, 0, sizeof(a));
    a.pixels = pixels; a.out = out; a.out_pitch = out_pitch; a.H = H; a.W = W; a.ch = ch;
    a.raw = (unsigned)H * (1u + (unsigned)W * ch);
    a.zlen = png_zlen(a.raw);
    a.fsize = (unsigned)png_file_bytes(H, W, ch);
    const unsigned fpad = (a.fsize + 15u) & ~15u;
    const size_t smem = (size_t)fpad + (256 + kPngThreads) * sizeof(uint32_t);
    if (smem > 227 * 1024 || 1ull + (unsigned long long)W * ch >= kStoredMax) {
        set_error("png_pack: a %d x %d x %d image (%u byte file) does not fit shared memory", H, W, ch, a.fsize);
        return MSL_ERR_UNSUPPORTED;
    }
    const unsigned clen = 4 + a.zlen;
    a.K = (clen + kPngThreads - 1) / kPngThreads;
    uint32_t m = xpow8(a.K);
    for (int l = 0; l < 10; ++l) { a.M[l] = m; m = multmodp(m, m); }
    a.crc_init_term = multmodp(xpow8(clen), 0xffffffffu);
    // signature, IHDR chunk, IDAT length + type
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t* h = a.head;
    memcpy(h, sig, 8);
    auto be32 = [](uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; };
    be32(h + 8, 13);
    memcpy(h + 12, "IHDR", 4);
    be32(h + 16, (uint32_t)W); be32(h + 20, (uint32_t)H);
    h[24] = 8; h[25] = ch == 4 ? 6 : 0; h[26] = 0; h[27] = 0; h[28] = 0;
    be32(h + 29, crc32_host(h + 12, 17));
    be32(h + 33, a.zlen);
    memcpy(h + 37, "IDAT", 4);
    ProfScope prof(K_PNG_PACK, stream);
    MSL_CUDA_CHECK(cudaFuncSetAttribute(png_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    png_pack_kernel<<<n, kPngThreads, smem, stream>>>(a);
    MSL_LAUNCH_CHECK("png_pack_kernel");
    return MSL_OK;
}

}  // namespace msl
