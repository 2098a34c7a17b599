// vi_api.cu -- C ABI (include/vi_b200.h) over the fused inspection kernel.
// Host side: context, parameter normalisation (the reference's widget rules),
// launch, the pipelined host-buffer batch call and the per-unit compat calls.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "vi_unit.cuh"
#include "vi_ingest.cuh"

using namespace vi;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(VI_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));  \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t n = 0;
    int ensure(size_t bytes) {
        if (bytes <= n) return VI_OK;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail(VI_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        n = bytes;
        return VI_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

constexpr int kHostSlots = 3;      // chunks in flight in the host-buffer batch call (upload / kernel / download overlap)

struct GridState {
    std::vector<int4> rects;
    std::vector<long long> off;     // n+1
    long long unit_px = 0;
    int wmax = 0, hmax = 0;
    DevBuf d_rects, d_off;
    SmemPlan plan{};
};

struct vi_ctx {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    int smem_static = 0;
    GridState grid;                 // user grid
    GridState one;                  // single-rect grid of the compat calls
    std::vector<vi_excl> excl;
    DevBuf d_excl;
    DevBuf d_refc;
    bool has_refc = false;
    int is_reference = 0;
    DevBuf scratch;
    long long scratch_stride = 0;
    long long scratch_f32_off = 0;
    long long scratch_rank_off = 0;
    // compat / host-batch staging
    DevBuf st_in, st_aux, st_out, st_out2, st_rec, st_stats, st_lab;
    DevBuf hb_frames[kHostSlots], hb_seg[kHostSlots], hb_def[kHostSlots], hb_rec[kHostSlots];
    cudaStream_t streams[kHostSlots] = {};
    int smem_set = 0;
    long long* prof = nullptr;
    long long* seg_stats = nullptr;
};

extern "C" const char* vi_last_error(void) { return g_err.c_str(); }
extern "C" int vi_version(void) { return 100; }

extern "C" void vi_params_default(vi_params* p) {
    if (!p) return;
    p->seg_method = 0; p->gaussian_blur = 3; p->morph_kernel = 3; p->adapt_block = 51; p->adapt_C = 10;
    p->defect_method = 0; p->threshold = 24; p->min_area = 20; p->erode_px = 6; p->median_ksize = 21;
    p->max_area_frac = 0.98;
}

extern "C" int vi_ctx_create(int device, vi_ctx** out) {
    if (!out) return fail(VI_ERR_ARG, "vi_ctx_create: out is null");
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(VI_ERR_ARG, "vi_ctx_create: device %d of %d", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    vi_ctx* c = new vi_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, vi_unit_kernel<false>));
    c->smem_static = ((int)fa.sharedSizeBytes + 15) & ~15;
    for (auto& s : c->streams) {
        cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail(VI_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    }
    *out = c;
    return VI_OK;
}

extern "C" void vi_ctx_destroy(vi_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (GridState* g : {&c->grid, &c->one}) { g->d_rects.release(); g->d_off.release(); }
    for (DevBuf* b : {&c->d_excl, &c->d_refc, &c->scratch, &c->st_in, &c->st_aux, &c->st_out, &c->st_out2, &c->st_rec,
                      &c->st_stats, &c->st_lab})
        b->release();
    for (int i = 0; i < kHostSlots; ++i) {
        c->hb_frames[i].release(); c->hb_seg[i].release(); c->hb_def[i].release(); c->hb_rec[i].release();
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    }
    delete c;
}

// Table uploads come from pageable host memory: cudaMemcpy may return while the DMA to the device is still in
// flight, and the kernels run on non-blocking streams that the legacy stream does not order against -- so the
// upload is either issued on the stream that launches next (`st`) or followed by a device synchronise.
static int set_grid_state(vi_ctx* c, GridState& gs, const int32_t* r, int n, cudaStream_t st = nullptr) {
    if (!r || n <= 0) return fail(VI_ERR_ARG, "grid: need at least one rect");
    gs.rects.resize(n);
    gs.off.assign(n + 1, 0);
    gs.wmax = gs.hmax = 0;
    for (int i = 0; i < n; ++i) {
        int x = r[4 * i], y = r[4 * i + 1], w = r[4 * i + 2], h = r[4 * i + 3];
        if (w <= 0 || h <= 0 || x < 0 || y < 0) return fail(VI_ERR_ARG, "grid: rect %d = (%d,%d,%d,%d) is invalid", i, x, y, w, h);
        if (w > 65535 || h > 65535) return fail(VI_ERR_TOO_LARGE, "grid: rect %d is %dx%d", i, w, h);
        gs.rects[i] = make_int4(x, y, w, h);
        gs.off[i + 1] = gs.off[i] + (long long)w * h;
        gs.wmax = std::max(gs.wmax, w);
        gs.hmax = std::max(gs.hmax, h);
    }
    gs.unit_px = gs.off[n];
    int gray_need = 0, words_need = 0;
    for (int i = 0; i < n; ++i) {
        const Geom g = make_geom(gs.rects[i].z, gs.rects[i].w);
        gray_need = std::max(gray_need, g.gp * g.h);
        words_need = std::max(words_need, g.nwords);
    }
    if (!make_plan(gs.wmax, gs.hmax, c->smem_optin, c->smem_static, &gs.plan, gray_need, words_need))
        return fail(VI_ERR_TOO_LARGE, "grid: its largest units (up to %d wide, %d high, %d crop bytes) do not fit the %d-byte "
                    "shared-memory-resident path", gs.wmax, gs.hmax, gray_need, c->smem_optin);
    int rc;
    if ((rc = gs.d_rects.ensure(sizeof(int4) * n))) return rc;
    if ((rc = gs.d_off.ensure(sizeof(long long) * (n + 1)))) return rc;
    if (st) {
        CU(cudaMemcpyAsync(gs.d_rects.p, gs.rects.data(), sizeof(int4) * n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(gs.d_off.p, gs.off.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
    } else {
        CU(cudaMemcpy(gs.d_rects.p, gs.rects.data(), sizeof(int4) * n, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(gs.d_off.p, gs.off.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice));
        CU(cudaDeviceSynchronize());
    }
    return VI_OK;
}

extern "C" int vi_set_grid(vi_ctx* c, const int32_t* rects, int n) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    CU(cudaSetDevice(c->device));
    c->has_refc = false;     // grid changed: reference centroids no longer valid (indexing_ui.py:2196-2200)
    return set_grid_state(c, c->grid, rects, n);
}

extern "C" int vi_set_exclusions(vi_ctx* c, const vi_excl* e, int n) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    if (n < 0 || (n > 0 && !e)) return fail(VI_ERR_ARG, "exclusions: bad arguments");
    if (n > 4096) return fail(VI_ERR_ARG, "exclusions: %d is too many", n);
    CU(cudaSetDevice(c->device));
    c->excl.assign(e, e + n);
    if (n) {
        int rc;
        if ((rc = c->d_excl.ensure(sizeof(vi_excl) * n))) return rc;
        CU(cudaMemcpy(c->d_excl.p, e, sizeof(vi_excl) * n, cudaMemcpyHostToDevice));
        CU(cudaDeviceSynchronize());
    }
    return VI_OK;
}

extern "C" int vi_set_ref_centroids(vi_ctx* c, const double* cxcy, int n_units, int is_reference) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    CU(cudaSetDevice(c->device));
    c->is_reference = is_reference ? 1 : 0;
    if (!cxcy) { c->has_refc = false; return VI_OK; }
    if (n_units != (int)c->grid.rects.size()) return fail(VI_ERR_ARG, "ref centroids: %d entries for %zu units", n_units, c->grid.rects.size());
    int rc;
    if ((rc = c->d_refc.ensure(sizeof(double) * 2 * n_units))) return rc;
    CU(cudaMemcpy(c->d_refc.p, cxcy, sizeof(double) * 2 * n_units, cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    c->has_refc = true;
    return VI_OK;
}

// Diagnostics: div_by_rcp against the IEEE divide on pseudo-random operand pairs drawn like
// the Otsu recurrence's (numerator in [0, 256), divisor in (2^-24, 1]), plus divisors with
// long runs of one bits in the significand.
__global__ void fastdiv_check_kernel(unsigned long long seed, long long per_thread, unsigned long long* bad) {
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long nbad = 0;
    for (long long i = 0; i < per_thread; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        unsigned long long m1 = s & 0x000fffffffffffffull;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        unsigned long long m2 = s & 0x000fffffffffffffull;
        const int kind = (int)((s >> 52) & 7);
        if (kind == 0) m2 |= 0x000fffffffff0000ull;                 // significand close to all ones
        if (kind == 1) m2 = 0x000fffffffffffffull - (m2 & 0xff);
        if (kind == 2) m1 &= 0x000ff00000000000ull;                 // short numerators (small integers / sums)
        const int eb = 1023 - (int)((s >> 56) % 24);                // divisor in (2^-24, 1]
        const int en = 1023 + 7 - (int)((s >> 60) % 12);            // numerator in [2^-4, 256)
        const double b = __longlong_as_double(((long long)eb << 52) | (long long)m2);
        const double n = __longlong_as_double(((long long)en << 52) | (long long)m1);
        const double r = __ddiv_rn(1.0, b);
        const double q = vi::div_by_rcp(n, b, r), ref = __ddiv_rn(n, b);
        if (__double_as_longlong(q) != __double_as_longlong(ref)) ++nbad;
    }
    if (nbad) atomicAdd(bad, nbad);
}

extern "C" int vi_debug_fastdiv_check(vi_ctx* c, long long n_samples, unsigned long long seed, long long* mismatches) {
    if (!c || !mismatches) return fail(VI_ERR_ARG, "vi_debug_fastdiv_check: null");
    CU(cudaSetDevice(c->device));
    unsigned long long* d_bad = nullptr;
    CU(cudaMalloc(&d_bad, 8));
    CU(cudaMemset(d_bad, 0, 8));
    const int blocks = 4 * c->sm_count, threads = 256;
    long long per = (n_samples + (long long)blocks * threads - 1) / ((long long)blocks * threads);
    fastdiv_check_kernel<<<blocks, threads>>>(seed, per, d_bad);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d_bad, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    if (e != cudaSuccess) return fail(VI_ERR_CUDA, "fastdiv check: %s", cudaGetErrorString(e));
    *mismatches = (long long)h;
    return VI_OK;
}

extern "C" int vi_debug_set_profile(vi_ctx* c, long long* d_cycles) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    c->prof = d_cycles;
    return VI_OK;
}

extern "C" int vi_set_seg_stats_output(vi_ctx* c, int64_t* d_stats) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    c->seg_stats = (long long*)d_stats;
    return VI_OK;
}

extern "C" int64_t vi_host_upload_bytes(vi_ctx* c, int n_images, int64_t row_pitch) {
    if (!c) return 0;
    std::vector<std::pair<int, int>> iv;
    for (const int4& r : c->grid.rects) iv.emplace_back(r.y, r.y + r.w);
    std::sort(iv.begin(), iv.end());
    long long rows = 0;
    int end = -1;
    for (const auto& p : iv) {
        int a0 = std::max(p.first, end);
        if (p.second > a0) rows += p.second - a0;
        end = std::max(end, p.second);
    }
    return (int64_t)rows * row_pitch * n_images;
}

extern "C" int64_t vi_unit_pixels(vi_ctx* c) { return c ? c->grid.unit_px : 0; }

extern "C" int vi_unit_offsets(vi_ctx* c, int64_t* out) {
    if (!c || !out) return fail(VI_ERR_ARG, "vi_unit_offsets: null");
    for (size_t i = 0; i < c->grid.off.size(); ++i) out[i] = c->grid.off[i];
    return VI_OK;
}

// ---------------------------------------------------------------------------
// parameter normalisation -- the reference's rules, on the host
// ---------------------------------------------------------------------------
static void gaussian_taps_q8(int k, int* q) {
    // OpenCV's 8.8 fixed-point kernel for uint8 GaussianBlur(k, sigma=0)
    static const double small3[] = {0.25, 0.5, 0.25};
    static const double small5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
    static const double small7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
    std::vector<double> kern(k);
    if (k == 1) kern[0] = 1.0;
    else if (k == 3) std::copy(small3, small3 + 3, kern.begin());
    else if (k == 5) std::copy(small5, small5 + 5, kern.begin());
    else if (k == 7) std::copy(small7, small7 + 7, kern.begin());
    else {
        double sigma = ((k - 1) * 0.5 - 1) * 0.3 + 0.8;
        double s2 = -0.5 / (sigma * sigma);
        double tot = 0;
        for (int i = 0; i < k; ++i) { double x = i - (k - 1) * 0.5; kern[i] = std::exp(s2 * x * x); tot += kern[i]; }
        double inv = 1.0 / tot;
        for (int i = 0; i < k; ++i) kern[i] *= inv;
    }
    double err = 0;
    int acc = 0;
    for (int i = 0; i < k / 2; ++i) {
        double adj = kern[i] * 256.0 + err;
        int v = (int)std::nearbyint(adj);      // cvRound: half to even
        err = adj - v;
        q[i] = q[k - 1 - i] = v;
        acc += v;
    }
    q[k / 2] = 256 - 2 * acc;
}

// cv2.getGaussianKernel(bs, 0, CV_32F): the float32 taps of adaptiveThreshold's Gaussian mean
// (hard-coded tables up to 9 taps in OpenCV 4.13, else exp(-x^2 / 2 sigma^2) normalised in double).
static void gaussian_taps_f32(int k, float* out) {
    static const double s3[] = {0.25, 0.5, 0.25};
    static const double s5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
    static const double s7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
    static const double s9[] = {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256, 51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256};
    const double* small = k == 3 ? s3 : k == 5 ? s5 : k == 7 ? s7 : k == 9 ? s9 : nullptr;
    if (k == 1) { out[0] = 1.0f; return; }
    if (small) { for (int i = 0; i < k; ++i) out[i] = (float)small[i]; return; }
    std::vector<double> kern(k);
    const double sigma = ((k - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double s2 = -0.5 / (sigma * sigma);
    double tot = 0;
    for (int i = 0; i < k; ++i) { const double x = i - (k - 1) * 0.5; kern[i] = std::exp(s2 * x * x); tot += kern[i]; }
    const double inv = 1.0 / tot;
    for (int i = 0; i < k; ++i) out[i] = (float)(kern[i] * inv);
}

extern "C" int vi_debug_adaptive_taps(int block_size, float* out) {
    if (!out || block_size < 1 || block_size > kMaxAdapt || block_size % 2 == 0) return fail(VI_ERR_ARG, "vi_debug_adaptive_taps: block size %d", block_size);
    gaussian_taps_f32(block_size, out);
    return VI_OK;
}

static void ellipse_spans(int k, signed char* lo, signed char* hi) {
    // cv2.getStructuringElement(MORPH_ELLIPSE,(k,k)) as per-row offset spans
    int r = k / 2, c = k / 2;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < k; ++i) {
        int dy = i - r;
        int j1 = 0, j2 = 0;
        if (std::abs(dy) <= r) {
            int dx = (int)std::nearbyint(c * std::sqrt((r * r - dy * dy) * inv_r2));
            j1 = std::max(c - dx, 0);
            j2 = std::min(c + dx + 1, k);
        }
        lo[i] = (signed char)(j1 - c);
        hi[i] = (signed char)(j2 - 1 - c);
    }
}

static int fill_params(KArgs& a, const vi_params* p) {
    if (!p) return fail(VI_ERR_ARG, "params is null");
    a.p = *p;
    if (p->seg_method != 0 && p->seg_method != 1) return fail(VI_ERR_ARG, "seg_method %d (0 otsu, 1 adaptive)", p->seg_method);
    a.adapt_bs = 0;
    if (p->seg_method == 1) {
        const int bs = std::max(3, p->adapt_block | 1);                                                     // segmentation.py:84
        if (bs > kMaxAdapt) return fail(VI_ERR_ARG, "adapt_block %d: block wider than %d", p->adapt_block, kMaxAdapt);
        a.adapt_bs = bs;
        gaussian_taps_f32(bs, a.ataps);
    }
    if (p->defect_method != 0 && p->defect_method != 1) return fail(VI_ERR_ARG, "defect_method %d (0 threshold, 1 canny)", p->defect_method);
    if (p->median_ksize != 21) return fail(VI_ERR_UNSUPPORTED, "median_ksize must be 21 (indexing_ui.py:1522)");
    if (p->threshold < 0 || p->threshold > 255) return fail(VI_ERR_ARG, "threshold %d outside 0..255", p->threshold);
    if (p->erode_px < 0 || p->min_area < 0) return fail(VI_ERR_ARG, "erode_px / min_area must be >= 0");
    a.canny_low = std::max(1, p->threshold / 2);                                                          // indexing_ui.py:1537
    a.canny_high = std::max(2, p->threshold);
    if (a.canny_low > a.canny_high) std::swap(a.canny_low, a.canny_high);
    int k = 0;
    if (p->gaussian_blur > 0) k = (p->gaussian_blur % 2 == 1) ? p->gaussian_blur : p->gaussian_blur + 1;   // segmentation.py:79
    if (k == 1) k = 0;                                     // a 1x1 Gaussian is the identity
    if (k > kMaxTaps) return fail(VI_ERR_ARG, "gaussian_blur %d: kernel wider than %d", p->gaussian_blur, kMaxTaps);
    a.blur_k = k;
    memset(a.taps, 0, sizeof a.taps);
    if (k >= 3) gaussian_taps_q8(k, a.taps);
    int mk = p->morph_kernel > 0 ? std::max(1, p->morph_kernel) : 0;                                   // segmentation.py:91-92
    if (mk == 1) mk = 0;                                   // a 1x1 element is the identity
    if (mk > kMaxSE) return fail(VI_ERR_ARG, "morph_kernel %d: element wider than %d", p->morph_kernel, kMaxSE);
    a.se_k = mk;
    memset(a.se_lo, 0, sizeof a.se_lo);
    memset(a.se_hi, 0, sizeof a.se_hi);
    if (mk > 0 && mk != 3) ellipse_spans(mk, a.se_lo, a.se_hi);
    return VI_OK;
}

static int ensure_scratch(vi_ctx* c, int wmax, int hmax, int nblocks, bool f32_plane) {
    long long px = (long long)wmax * hmax;
    long long capg = (long long)hmax * (wmax / 2 + 1);
    long long px4 = (long long)((wmax + 3) & ~3) * hmax;
    long long stride = ((px * 2 + 15) & ~15ll) + ((px4 + 15) & ~15ll) + (long long)ccl_ws_bytes((int)capg, hmax) + 256;
    stride = (stride + 255) & ~255ll;
    c->scratch_rank_off = stride;
    stride += (rank_scratch_bytes(wmax, hmax) + 255) & ~255ll;
    c->scratch_f32_off = stride;
    if (f32_plane) stride += (px * 4 + 255) & ~255ll;       // float plane of the adaptive mean, only when asked for
    c->scratch_stride = stride;
    return c->scratch.ensure((size_t)stride * nblocks);
}

// `slot` selects a private copy of the per-CTA scratch so launches on the internal
// streams never share it.
static int launch_units(vi_ctx* c, KArgs& a, const GridState& gs, cudaStream_t stream, int slot = 0) {
    const long long n_total = (long long)a.n_images * a.n_units;
    if (n_total <= 0) return VI_OK;
    if (n_total > 0x7fffffffll) return fail(VI_ERR_ARG, "too many units in one call");
    int nblocks = (int)std::min<long long>(n_total, c->sm_count);
    int rc;
    if ((rc = ensure_scratch(c, gs.wmax, gs.hmax, kHostSlots * c->sm_count, a.p.seg_method == 1))) return rc;
    a.scratch = (uint8_t*)c->scratch.p + (size_t)slot * c->sm_count * c->scratch_stride;
    a.scratch_stride = c->scratch_stride;
    a.scratch_f32_off = c->scratch_f32_off;
    a.scratch_rank_off = c->scratch_rank_off;
    a.wmax = gs.wmax; a.hmax = gs.hmax;
    a.plan = gs.plan;
    a.prof = (&gs == &c->grid) ? c->prof : nullptr;
    if (&gs != &c->grid) a.seg_stats = nullptr;
    if (c->smem_set < gs.plan.total) {
        CU(cudaFuncSetAttribute(vi_unit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - c->smem_static));
        CU(cudaFuncSetAttribute(vi_unit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - c->smem_static));
        c->smem_set = c->smem_optin - c->smem_static;
    }
    if (gs.plan.total > c->smem_set) return fail(VI_ERR_TOO_LARGE, "shared-memory plan %d > %d", gs.plan.total, c->smem_set);
    if (a.prof) vi_unit_kernel<true><<<nblocks, kThreads, gs.plan.total, stream>>>(a);       // diagnostics build: phase timers
    else vi_unit_kernel<false><<<nblocks, kThreads, gs.plan.total, stream>>>(a);
    CU(cudaGetLastError());
    return VI_OK;
}

static int check_frames(const GridState& gs, int n_images, int W, int H, int64_t row_pitch, int64_t image_stride) {
    if (n_images <= 0 || W <= 0 || H <= 0) return fail(VI_ERR_ARG, "frames: n_images=%d W=%d H=%d", n_images, W, H);
    if (row_pitch < W || (n_images > 1 && image_stride < row_pitch * (int64_t)H))
        return fail(VI_ERR_ARG, "frames: pitch/stride smaller than the image");
    for (size_t i = 0; i < gs.rects.size(); ++i) {
        const int4& r = gs.rects[i];
        // the reference pads (QImage.copy) or clips (QPixmap.copy) out-of-frame rects, inconsistently; reject them
        if (r.x + r.z > W || r.y + r.w > H) return fail(VI_ERR_ARG, "grid: rect %zu (%d,%d,%d,%d) leaves the %dx%d frame", i, r.x, r.y, r.z, r.w, W, H);
    }
    return VI_OK;
}

static void base_args(vi_ctx* c, KArgs& a, const GridState& gs) {
    memset(&a, 0, sizeof a);
    a.rects = (const int4*)gs.d_rects.p;
    a.n_units = (int)gs.rects.size();
    a.unit_off = (const long long*)gs.d_off.p;
    a.unit_px = gs.unit_px;
}

extern "C" int vi_inspect_batch(vi_ctx* c, const uint8_t* d_frames, int n_images, int W, int H, int64_t row_pitch,
                                int64_t image_stride, const vi_params* params, uint8_t* d_seg, uint8_t* d_def,
                                int32_t* d_labels, vi_unit_record* d_rec, void* stream) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    if (c->grid.rects.empty()) return fail(VI_ERR_ARG, "vi_inspect_batch: no grid set");
    if (!d_frames || !d_rec) return fail(VI_ERR_ARG, "vi_inspect_batch: frames / records pointer is null");
    CU(cudaSetDevice(c->device));
    int rc;
    if ((rc = check_frames(c->grid, n_images, W, H, row_pitch, image_stride))) return rc;
    KArgs a;
    base_args(c, a, c->grid);
    if ((rc = fill_params(a, params))) return rc;
    a.frames = d_frames; a.n_images = n_images; a.W = W; a.H = H; a.row_pitch = row_pitch; a.image_stride = image_stride;
    a.excl = (const vi_excl*)c->d_excl.p; a.n_excl = (int)c->excl.size();
    a.refc = c->has_refc ? (const double*)c->d_refc.p : nullptr;
    a.is_reference = c->is_reference;
    a.seg_out = d_seg; a.def_out = d_def; a.labels_out = d_labels; a.rec = d_rec;
    a.mode = MODE_FULL;
    a.seg_stats = c->seg_stats;
    return launch_units(c, a, c->grid, (cudaStream_t)stream);
}

extern "C" int vi_inspect_batch_host(vi_ctx* c, const uint8_t* h_frames, int n_images, int W, int H, int64_t row_pitch,
                                     int64_t image_stride, const vi_params* params, uint8_t* h_seg, uint8_t* h_def,
                                     vi_unit_record* h_rec) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    if (c->grid.rects.empty()) return fail(VI_ERR_ARG, "vi_inspect_batch_host: no grid set");
    if (!h_frames || !h_rec) return fail(VI_ERR_ARG, "vi_inspect_batch_host: frames / records pointer is null");
    CU(cudaSetDevice(c->device));
    int rc;
    if ((rc = check_frames(c->grid, n_images, W, H, row_pitch, image_stride))) return rc;
    const int n_units = (int)c->grid.rects.size();
    const long long upx = c->grid.unit_px;
    // Chunks of about two units per SM, kHostSlots of them in flight.  The call is PCIe-bound (the two byte masks
    // down, the covered frame rows up: measured ~76 GB/s aggregate over both directions, chunk sizes from 2 to 13
    // images within 10 % of each other), so all that matters is that the copy engines never idle.
    // VI_HOST_CHUNK overrides (images per chunk).
    int chunk = std::max(1, (2 * c->sm_count + n_units - 1) / n_units);
    if (const char* e = getenv("VI_HOST_CHUNK")) { int v = atoi(e); if (v > 0) chunk = v; }
    chunk = std::min(chunk, n_images);
    const size_t frame_bytes = (size_t)image_stride;
    for (int b = 0; b < kHostSlots; ++b) {
        if ((rc = c->hb_frames[b].ensure(frame_bytes * chunk))) return rc;
        if ((rc = c->hb_seg[b].ensure((size_t)upx * chunk))) return rc;
        if ((rc = c->hb_def[b].ensure((size_t)upx * chunk))) return rc;
        if ((rc = c->hb_rec[b].ensure(sizeof(vi_unit_record) * (size_t)n_units * chunk))) return rc;
    }
    // merged [y0, y1) row intervals covered by the grid
    std::vector<std::pair<int, int>> rows;
    {
        std::vector<std::pair<int, int>> iv;
        for (const int4& r : c->grid.rects) iv.emplace_back(r.y, r.y + r.w);
        std::sort(iv.begin(), iv.end());
        for (const auto& p : iv) {
            if (!rows.empty() && p.first <= rows.back().second) rows.back().second = std::max(rows.back().second, p.second);
            else rows.push_back(p);
        }
    }
    int slot = 0;
    for (int i0 = 0; i0 < n_images; i0 += chunk, slot = (slot + 1) % kHostSlots) {
        const int n = std::min(chunk, n_images - i0);
        cudaStream_t st = c->streams[slot];
        // the slot's previous chunk (kHostSlots iterations ago) is ordered before this one on the same stream
        // upload only the frame rows some unit covers: one strided copy per merged row interval
        for (const auto& iv : rows) {
            const size_t off = (size_t)iv.first * row_pitch, bytes = (size_t)(iv.second - iv.first) * row_pitch;
            CU(cudaMemcpy2DAsync((uint8_t*)c->hb_frames[slot].p + off, frame_bytes, h_frames + (size_t)i0 * image_stride + off,
                                 (size_t)image_stride, bytes, n, cudaMemcpyHostToDevice, st));
        }
        KArgs a;
        base_args(c, a, c->grid);
        if ((rc = fill_params(a, params))) return rc;
        a.frames = (const uint8_t*)c->hb_frames[slot].p; a.n_images = n; a.W = W; a.H = H; a.row_pitch = row_pitch;
        a.image_stride = image_stride;
        a.excl = (const vi_excl*)c->d_excl.p; a.n_excl = (int)c->excl.size();
        a.refc = c->has_refc ? (const double*)c->d_refc.p : nullptr;
        a.is_reference = c->is_reference;
        a.seg_out = (uint8_t*)c->hb_seg[slot].p; a.def_out = (uint8_t*)c->hb_def[slot].p; a.rec = (vi_unit_record*)c->hb_rec[slot].p;
        a.mode = MODE_FULL;
        if ((rc = launch_units(c, a, c->grid, st, slot))) return rc;
        if (h_seg) CU(cudaMemcpyAsync(h_seg + (size_t)i0 * upx, c->hb_seg[slot].p, (size_t)upx * n, cudaMemcpyDeviceToHost, st));
        if (h_def) CU(cudaMemcpyAsync(h_def + (size_t)i0 * upx, c->hb_def[slot].p, (size_t)upx * n, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_rec + (size_t)i0 * n_units, c->hb_rec[slot].p, sizeof(vi_unit_record) * (size_t)n_units * n,
                           cudaMemcpyDeviceToHost, st));
    }
    for (int b = 0; b < kHostSlots; ++b) CU(cudaStreamSynchronize(c->streams[b]));
    // records carry chunk-local image indices: make them batch-global
    for (int i0 = 0; i0 < n_images; i0 += chunk) {
        const int n = std::min(chunk, n_images - i0);
        for (long long k = 0; k < (long long)n * n_units; ++k) h_rec[(size_t)i0 * n_units + k].image += i0;
    }
    return VI_OK;
}

// ---------------------------------------------------------------------------
// compat entry points: one unit = the whole (h, w) array, host pointers
// ---------------------------------------------------------------------------
static int compat_run(vi_ctx* c, int mode, const uint8_t* gray, const uint8_t* aux, int h, int w, const vi_params* params,
                      int erode_r, uint8_t* out_seg, uint8_t* out_def, int32_t* out_lab, vi_unit_record* out_rec,
                      long long* out_stats /*[8]*/) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    if (h <= 0 || w <= 0) return fail(VI_ERR_ARG, "array is %dx%d", h, w);
    CU(cudaSetDevice(c->device));
    int32_t rect[4] = {0, 0, w, h};
    int rc;
    cudaStream_t st = c->streams[0];
    if ((rc = set_grid_state(c, c->one, rect, 1, st))) return rc;
    const size_t px = (size_t)w * h;
    KArgs a;
    base_args(c, a, c->one);
    vi_params dflt;
    vi_params_default(&dflt);
    if ((rc = fill_params(a, params ? params : &dflt))) return rc;
    if ((rc = c->st_in.ensure(px + 16)) || (rc = c->st_aux.ensure(px + 16)) || (rc = c->st_out.ensure(px + 16)) ||
        (rc = c->st_out2.ensure(px + 16)) || (rc = c->st_rec.ensure(sizeof(vi_unit_record))) ||
        (rc = c->st_stats.ensure(64)) || (rc = c->st_lab.ensure(px * 4 + 16)))
        return rc;
    if (gray) CU(cudaMemcpyAsync(c->st_in.p, gray, px, cudaMemcpyHostToDevice, st));
    if (aux) CU(cudaMemcpyAsync(c->st_aux.p, aux, px, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->st_rec.p, 0, sizeof(vi_unit_record), st));
    CU(cudaMemsetAsync(c->st_stats.p, 0, 64, st));
    a.frames = (const uint8_t*)c->st_in.p; a.n_images = 1; a.W = w; a.H = h; a.row_pitch = w; a.image_stride = (long long)px;
    a.aux_mask = (const uint8_t*)c->st_aux.p;
    a.seg_out = (uint8_t*)c->st_out.p; a.def_out = (uint8_t*)c->st_out2.p;
    a.labels_out = out_lab ? (int32_t*)c->st_lab.p : nullptr;
    a.rec = (vi_unit_record*)c->st_rec.p;
    a.stats_out = (long long*)c->st_stats.p;
    a.mode = mode;
    a.erode_r = erode_r;
    if ((rc = launch_units(c, a, c->one, st))) return rc;
    if (out_seg) CU(cudaMemcpyAsync(out_seg, c->st_out.p, px, cudaMemcpyDeviceToHost, st));
    if (out_def) CU(cudaMemcpyAsync(out_def, c->st_out2.p, px, cudaMemcpyDeviceToHost, st));
    if (out_lab) CU(cudaMemcpyAsync(out_lab, c->st_lab.p, px * 4, cudaMemcpyDeviceToHost, st));
    if (out_rec) CU(cudaMemcpyAsync(out_rec, c->st_rec.p, sizeof(vi_unit_record), cudaMemcpyDeviceToHost, st));
    if (out_stats) CU(cudaMemcpyAsync(out_stats, c->st_stats.p, 64, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return VI_OK;
}

extern "C" int vi_segment_cell(vi_ctx* c, const uint8_t* gray, int h, int w, const vi_params* params, uint8_t* out_mask,
                               int32_t* out_otsu_t) {
    if (!gray || !out_mask) return fail(VI_ERR_ARG, "vi_segment_cell: null pointer");
    vi_unit_record rec;
    int rc = compat_run(c, MODE_SEG_ONLY, gray, nullptr, h, w, params, 0, out_mask, nullptr, nullptr, &rec, nullptr);
    if (rc == VI_OK && out_otsu_t) *out_otsu_t = rec.otsu_t;
    return rc;
}

extern "C" int vi_fill_internal_holes(vi_ctx* c, const uint8_t* mask, int h, int w, uint8_t* out_mask) {
    if (!mask || !out_mask) return fail(VI_ERR_ARG, "vi_fill_internal_holes: null pointer");
    return compat_run(c, MODE_FILL, nullptr, mask, h, w, nullptr, 0, out_mask, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int vi_mask_stats(vi_ctx* c, const uint8_t* mask, int h, int w, int64_t* area, int64_t* sum_x, int64_t* sum_y) {
    if (!mask) return fail(VI_ERR_ARG, "vi_mask_stats: null pointer");
    long long st[8];
    int rc = compat_run(c, MODE_STATS, nullptr, mask, h, w, nullptr, 0, nullptr, nullptr, nullptr, nullptr, st);
    if (rc) return rc;
    if (area) *area = st[0];
    if (sum_x) *sum_x = st[1];
    if (sum_y) *sum_y = st[2];
    return VI_OK;
}

extern "C" int vi_erode_square(vi_ctx* c, const uint8_t* mask, int h, int w, int r, uint8_t* out_mask) {
    if (!mask || !out_mask) return fail(VI_ERR_ARG, "vi_erode_square: null pointer");
    if (r < 0) return fail(VI_ERR_ARG, "vi_erode_square: r = %d", r);
    return compat_run(c, MODE_ERODE, nullptr, mask, h, w, nullptr, r, out_mask, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int vi_label_components(vi_ctx* c, const uint8_t* mask, int h, int w, int32_t* out_labels, int32_t* n_labels,
                                   int32_t* best_label, int64_t* best_area, int64_t* best_sum_x, int64_t* best_sum_y) {
    if (!mask) return fail(VI_ERR_ARG, "vi_label_components: null pointer");
    long long st[8];
    int rc = compat_run(c, MODE_LABEL, nullptr, mask, h, w, nullptr, 0, nullptr, nullptr, out_labels, nullptr, st);
    if (rc) return rc;
    if (n_labels) *n_labels = (int32_t)st[0];
    if (best_label) *best_label = (int32_t)st[1];
    if (best_area) *best_area = st[2];
    if (best_sum_x) *best_sum_x = st[3];
    if (best_sum_y) *best_sum_y = st[4];
    return VI_OK;
}

extern "C" int vi_detect_defects(vi_ctx* c, const uint8_t* gray, const uint8_t* seg_mask, int h, int w,
                                 const vi_params* params, uint8_t* out_mask, int32_t* found, vi_unit_record* out_rec) {
    if (!gray || !seg_mask || !out_mask) return fail(VI_ERR_ARG, "vi_detect_defects: null pointer");
    vi_unit_record rec;
    int rc = compat_run(c, MODE_DETECT, gray, seg_mask, h, w, params, 0, nullptr, out_mask, nullptr, &rec, nullptr);
    if (rc) return rc;
    if (found) *found = rec.n_kept > 0 ? 1 : 0;
    if (out_rec) *out_rec = rec;
    return VI_OK;
}

// ---------------------------------------------------------------------------
// frame ingest (device pointers, asynchronous on `stream`)
// ---------------------------------------------------------------------------
static int ingest_common(vi_ctx* c, int fmt, const void* d_src, int n_images, int W, int H, int64_t src_pitch,
                         int64_t src_stride, uint8_t* d_dst, int64_t dst_pitch, int64_t dst_stride, void* stream) {
    if (!c) return fail(VI_ERR_ARG, "ctx is null");
    if (!d_src || !d_dst) return fail(VI_ERR_ARG, "ingest: null pointer");
    const int bpp = fmt == 0 ? 4 : 2;
    if (n_images <= 0 || W <= 0 || H <= 0) return fail(VI_ERR_ARG, "ingest: n_images=%d W=%d H=%d", n_images, W, H);
    if (src_pitch < (int64_t)W * bpp || dst_pitch < W) return fail(VI_ERR_ARG, "ingest: pitch smaller than a row");
    if (n_images > 1 && (src_stride < src_pitch * H || dst_stride < dst_pitch * H)) return fail(VI_ERR_ARG, "ingest: stride smaller than an image");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint8_t* src = (const uint8_t*)d_src;
    const bool aligned = ((uintptr_t)src % 16 == 0) && ((uintptr_t)d_dst % 16 == 0) && src_pitch % 16 == 0 && src_stride % 16 == 0 &&
                         dst_pitch % 16 == 0 && dst_stride % 16 == 0 && W % 16 == 0;
    const int blocks = c->sm_count * 8;                 // 8 CTAs of 256 threads per SM: a full complement of warps
    if (aligned) {
        const long long n_items = (long long)n_images * H * (W / 16);
        if (fmt == 0) ingest_argb32_v16<<<blocks, kIngestThreads, 0, st>>>(src, src_pitch, src_stride, d_dst, dst_pitch, dst_stride, W, H, n_items);
        else ingest_gray16_v16<<<blocks, kIngestThreads, 0, st>>>(src, src_pitch, src_stride, d_dst, dst_pitch, dst_stride, W, H, n_items);
    } else {
        const long long n_px = (long long)n_images * H * W;
        if (fmt == 0) ingest_argb32_scalar<<<blocks, kIngestThreads, 0, st>>>(src, src_pitch, src_stride, d_dst, dst_pitch, dst_stride, W, H, n_px);
        else ingest_gray16_scalar<<<blocks, kIngestThreads, 0, st>>>(src, src_pitch, src_stride, d_dst, dst_pitch, dst_stride, W, H, n_px);
    }
    CU(cudaGetLastError());
    return VI_OK;
}

extern "C" int vi_ingest_argb32(vi_ctx* c, const uint8_t* d_bgra, int n_images, int W, int H, int64_t src_pitch,
                                int64_t src_stride, uint8_t* d_gray, int64_t dst_pitch, int64_t dst_stride, void* stream) {
    return ingest_common(c, 0, d_bgra, n_images, W, H, src_pitch, src_stride, d_gray, dst_pitch, dst_stride, stream);
}

extern "C" int vi_ingest_gray16(vi_ctx* c, const uint16_t* d_gray16, int n_images, int W, int H, int64_t src_pitch,
                                int64_t src_stride, uint8_t* d_gray, int64_t dst_pitch, int64_t dst_stride, void* stream) {
    return ingest_common(c, 1, d_gray16, n_images, W, H, src_pitch, src_stride, d_gray, dst_pitch, dst_stride, stream);
}
