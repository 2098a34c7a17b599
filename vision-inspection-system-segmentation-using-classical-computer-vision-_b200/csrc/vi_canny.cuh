// vi_canny.cuh -- the detector's 'canny' branch: cv2.Canny(gray, max(1, thr/2), max(2, thr))
// (indexing_ui.py:1536-1539; aperture 3, L1 gradient).  OpenCV's algorithm, integer throughout:
//   Sobel 3x3 with BORDER_REPLICATE at the crop edge; magnitude m = |gx| + |gy|, 0 outside the crop;
//   non-maximum suppression along the gradient direction quantised by tan 22.5 / tan 67.5 in 15-bit
//   fixed point (TG22 = 13573): horizontal m > left && m >= right, vertical m > up && m >= down,
//   diagonal strict on both sides, the diagonal chosen by the sign of gx ^ gy;
//   candidates need m > low, strong ones m > high;
//   hysteresis = every candidate 8-connected (through candidates) to a strong one.
// The hysteresis flood is the CTA's run-based labelling (vi_ccl.cuh) plus a per-component flag.
// oracle/restate.py: canny_edges is the numpy twin.
#pragma once
#include "vi_ccl.cuh"

namespace vi {

__device__ __forceinline__ void sobel_at(const uint8_t* gray, const Geom& g, int x, int y, int& gx, int& gy) {
    const int xl = max(x - 1, 0), xr = min(x + 1, g.w - 1);
    const uint8_t* r0 = gray + max(y - 1, 0) * g.gp;
    const uint8_t* r1 = gray + y * g.gp;
    const uint8_t* r2 = gray + min(y + 1, g.h - 1) * g.gp;
    const int a = r0[xl], b = r0[x], c = r0[xr], d = r1[xl], e = r1[xr], f = r2[xl], gg = r2[x], hh = r2[xr];
    gx = (c + 2 * e + hh) - (a + 2 * d + f);
    gy = (f + 2 * gg + hh) - (a + 2 * b + c);
}

__device__ __forceinline__ int sobel_mag_at(const uint8_t* gray, const Geom& g, int x, int y) {
    if ((unsigned)x >= (unsigned)g.w || (unsigned)y >= (unsigned)g.h) return 0;
    int gx, gy;
    sobel_at(gray, g, x, y, gx, gy);
    return abs(gx) + abs(gy);
}

// CAND = pixels that survive non-maximum suppression with m > low; STRONG = those with m > high.
VI_PHASE void canny_candidates(const uint8_t* gray, const Geom& g, int low, int high, unsigned* CAND, unsigned* STRONG) {
    const int lane = lane_id();
    for (int i = warp_id(); i < g.nwords; i += kWarps) {
        int y, c; word_rc(g, i, y, c);
        const int x = c * 32 + lane;
        bool cand = false, strong = false;
        if (x < g.w) {
            int gx, gy;
            sobel_at(gray, g, x, y, gx, gy);
            const int m = abs(gx) + abs(gy);
            if (m > low) {
                const int ax = abs(gx), ay15 = abs(gy) << 15;
                const int tg22x = ax * 13573;
                bool ok;
                if (ay15 < tg22x) {
                    ok = m > sobel_mag_at(gray, g, x - 1, y) && m >= sobel_mag_at(gray, g, x + 1, y);
                } else if (ay15 > tg22x + (ax << 16)) {
                    ok = m > sobel_mag_at(gray, g, x, y - 1) && m >= sobel_mag_at(gray, g, x, y + 1);
                } else {
                    const int s = (gx ^ gy) < 0 ? -1 : 1;
                    ok = m > sobel_mag_at(gray, g, x - s, y - 1) && m > sobel_mag_at(gray, g, x + s, y + 1);
                }
                cand = ok;
                strong = ok && m > high;
            }
        }
        const unsigned cb = __ballot_sync(kFull, cand), sb = __ballot_sync(kFull, strong);
        if (lane == 0) { CAND[i] = cb; STRONG[i] = sb; }
    }
}

// EDGES = the 8-components of CAND that hold a STRONG pixel.
VI_PHASE int canny_hysteresis(Cta& cs, const unsigned* CAND, const unsigned* STRONG, unsigned* EDGES, const Geom& g,
                                       const CclWs& ws_s, const CclWs& ws_g, CclWs& ws) {
    const int R = ccl_build(cs, CAND, g, true, false, ws_s, ws_g, ws, (PhaseTimerT<false>*)nullptr);
    for (int i = 1 + threadIdx.x; i <= R; i += kThreads) {
        const int y = ws.yy()[i], xs = ws.xs()[i], xe = ws.xe()[i];
        unsigned hit = 0;
        for (int c = xs >> 5; c <= (xe >> 5); ++c)
            hit |= STRONG[y * g.wpr + c] & bit_range(max(xs, c * 32) - c * 32, min(xe, c * 32 + 31) - c * 32);
        if (hit) ws.acc0()[ws.parent()[i]] = 1u;          // every writer stores the same value
    }
    cta_sync();
    const unsigned* acc0 = ws.acc0();
    ccl_paint(EDGES, nullptr, g, ws, [acc0](int root) { return acc0[root] != 0u; });
    cta_sync();
    return R;
}

}  // namespace vi
