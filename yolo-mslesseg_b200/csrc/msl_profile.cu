// Launch accounting and opt-in per-launch CUDA-event timing (used by bench.py for the roofline line).
// Counting is always on (one relaxed atomic per launch); event timing only between
// msl_profile_enable(1) and msl_profile_collect().  Events are recorded on the launching stream.
#include <atomic>
#include <mutex>
#include <vector>

#include "msl_common.cuh"
#include "msl_kernels.h"

namespace msl {

namespace {
std::atomic<unsigned long long> g_launches[K_NKIND];
std::atomic<bool> g_on{false};
struct Rec { int kind; cudaEvent_t e0, e1; cudaStream_t stream; };
std::mutex g_mu;
std::vector<Rec> g_recs;

const char* const kNames[K_NKIND] = {
    "enhance_slices_f32_none", "enhance_slices_f32_he", "enhance_slices_f32_clahe", "enhance_slices_f32_gc", "enhance_slices_f32_lt",
    "enhance_slices_u8_none", "enhance_slices_u8_he", "enhance_slices_u8_clahe", "enhance_slices_u8_gc", "enhance_slices_u8_lt",
    "init_stats", "plane_stats_f32", "lesion_flags", "norm_scatter",
    "recon_fill", "recon_slot_map", "recon_gather", "consensus_eval", "confusion_counts", "enhance_dense",
    "combine_predictions", "slice_counts", "bgr_to_gray", "png_pack", "nonzero_flags", "enhance_dense_tables",
    "deflate", "deflate_scan", "deflate_pack", "inflate", "png_unfilter", "nifti_convert", "checksum", "contours", "stage_slices",
};
}  // namespace

ProfScope::ProfScope(int kind, cudaStream_t stream) : kind_(kind), stream_(stream), e0_(nullptr), e1_(nullptr) {
    g_launches[kind].fetch_add(1, std::memory_order_relaxed);
    if (g_on.load(std::memory_order_relaxed)) {
        if (cudaEventCreate(&e0_) == cudaSuccess && cudaEventCreate(&e1_) == cudaSuccess) cudaEventRecord(e0_, stream_);
        else { e0_ = nullptr; e1_ = nullptr; }
    }
}

ProfScope::~ProfScope() {
    if (e0_ && e1_) {
        cudaEventRecord(e1_, stream_);
        std::lock_guard<std::mutex> lk(g_mu);
        g_recs.push_back({kind_, e0_, e1_, stream_});
    }
}

}  // namespace msl

using namespace msl;

extern "C" {

int msl_kernel_kinds(void) { return K_NKIND; }

const char* msl_kernel_name(int kind) { return (kind >= 0 && kind < K_NKIND) ? kNames[kind] : ""; }

unsigned long long msl_kernel_launches(unsigned long long* per_kind) {
    unsigned long long tot = 0;
    for (int k = 0; k < K_NKIND; ++k) {
        unsigned long long n = g_launches[k].load(std::memory_order_relaxed);
        if (per_kind) per_kind[k] = n;
        tot += n;
    }
    return tot;
}

int msl_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (on) {
        for (auto& r : g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        g_recs.clear();
    }
    g_on.store(on != 0);
    return MSL_OK;
}

int msl_profile_timeline(int cap, int* kind, unsigned long long* stream, double* start_ms, double* end_ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    for (auto& r : g_recs) {
        if (n >= cap) break;
        float a = 0.f, b = 0.f;
        if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&a, g_recs[0].e0, r.e0) != cudaSuccess ||
            cudaEventElapsedTime(&b, g_recs[0].e0, r.e1) != cudaSuccess) {
            set_error("profile event failed"); return -1;
        }
        kind[n] = r.kind; stream[n] = (unsigned long long)reinterpret_cast<uintptr_t>(r.stream); start_ms[n] = a; end_ms[n] = b;
        ++n;
    }
    return n;
}

int msl_profile_collect(double* ms_per_kind, unsigned long long* n_per_kind) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_on.store(false);
    for (int k = 0; k < K_NKIND; ++k) {
        if (ms_per_kind) ms_per_kind[k] = 0.0;
        if (n_per_kind) n_per_kind[k] = 0;
    }
    int rc = MSL_OK;
    for (auto& r : g_recs) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (e != cudaSuccess) { set_error("profile event failed: %s", cudaGetErrorString(e)); rc = MSL_ERR_CUDA; }
        else {
            if (ms_per_kind) ms_per_kind[r.kind] += (double)ms;
            if (n_per_kind) n_per_kind[r.kind] += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_recs.clear();
    return rc;
}

}  // extern "C"
