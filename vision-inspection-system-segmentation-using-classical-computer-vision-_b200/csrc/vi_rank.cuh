// vi_rank.cuh -- P11: |gray - median21(gray)| > thr inside the ROI, without ever
// forming the median (cv2.medianBlur(gray, 21) + absdiff + threshold,
// indexing_ui.py:1522-1527; BORDER_REPLICATE).
//
//   med >  g+thr   <=>  #(window <= g+thr)   <= 220
//   med <= g-thr-1 <=>  #(window <= g-thr-1) >= 221          (441-pixel window)
//
// 1. Six unit-wide levels v_0 <= .. <= v_5 (around the two Otsu class medians).
//    C_k(p) = #(window(p) <= v_k) is a 21x21 box sum of per-pixel indicators, kept
//    for all six levels at once in two words of three 10-bit fields (441 < 1024).
// 2. C_k is evaluated exactly only on a lattice: one point per 3x3 cell (its
//    centre).  Moving the window by one pixel swaps one 21-pixel row or column, so
//    |C_k(p) - C_k(q)| <= 21 * L1(p, q) <= 42 inside a cell.  Hence for every pixel
//    of the cell   C_k >= 221 if C_k(centre) >= 263   and   C_k <= 220 if
//    C_k(centre) <= 178,   which brackets the median of every pixel of the cell:
//    LO < med <= HI with LO, HI taken from the level list (or -1 / 255).
// 3. Four byte thresholds per cell decide a pixel: defect for sure (g <= LO-thr or
//    g >= HI+thr+1), clean for sure (HI-thr <= g <= LO+thr+1), else "ambiguous".
//    A cell whose 9 pixels' min and max are both in the clean range is finished
//    (nearly all are); the few "dirty" cells are listed and classified per pixel,
//    and ambiguous ROI pixels get an exact rank count at their own two pivots.
//    Exact for any level set and any image
//    (oracle/restate.py: residual_mask_lattice is the numpy twin).
//
// Box sums are separable.  V: one thread per column accumulates 3-row block sums
// (a lattice row's 21-row window is exactly 7 blocks); a ring of packed block sums
// gives the sliding 7-block sum.  H: one warp per lattice row, one lane per cell:
// 3-column triple sums (10 replicated columns each side), a warp scan, and the
// cell's 21-column window is P[i+6] - P[i-1].
#pragma once
#include "vi_device.cuh"

namespace vi {

constexpr int kCell = 3;
constexpr int kLatBand = 16;             // lattice rows per band (= warps per CTA)
constexpr int kCellsPerPass = 26;        // 32 triples per pass, 6 of them look-ahead
constexpr int kPassBatch = 5;            // passes interleaved per lattice row (5 x 26 cells = 390 px)
constexpr unsigned kFld = 0x00300C03u;   // 2-bit block-sum fields at the 10-bit field positions
constexpr unsigned kFlag = 0x20080200u;  // bit 9 of each 10-bit field
constexpr unsigned kGe263 = 249u | (249u << 10) | (249u << 20);   // field + 249 >= 512  <=>  field >= 263
constexpr unsigned kGe179 = 333u | (333u << 10) | (333u << 20);   // field + 333 >= 512  <=>  field >= 179

struct RankWs {
    unsigned* lut0;     // [256] indicators of levels 0..2 (10-bit fields)
    unsigned* lut1;     // [256] indicators of levels 3..5
    unsigned* ring;     // [8][pitch] packed 3-row block sums
    uint2* cs;          // [kLatBand][pitch] 7-block column sums
    uint2* dirty;       // [dirty_cap] (cell position, thresholds); more are classified inline
    unsigned* exact;    // [exact_cap] ambiguous pixels (y << 16 | x); more are counted inline
    int dirty_cap, exact_cap;
    int* counters;      // [0] dirty cells, [1] ambiguous pixels listed, [2] ambiguous pixels total
    int pitch;
};

__host__ __device__ inline int rank_ws_bytes(int w) {
    const int pitch = (w + 3) & ~3;
    return 2 * 256 * 4 + 8 * pitch * 4 + kLatBand * pitch * 8 + 64;
}

// The two lists live in two mask buffers that are idle during this stage.
__device__ inline RankWs rank_ws_carve(unsigned char* base, int w, unsigned* listA, unsigned* listB, int mask_bytes) {
    RankWs r;
    r.pitch = (w + 3) & ~3;
    r.cs = reinterpret_cast<uint2*>(base); base += kLatBand * r.pitch * 8;
    r.dirty = reinterpret_cast<uint2*>(listA); r.dirty_cap = mask_bytes / 8;
    r.exact = listB; r.exact_cap = mask_bytes / 4;
    r.lut0 = reinterpret_cast<unsigned*>(base); base += 256 * 4;
    r.lut1 = reinterpret_cast<unsigned*>(base); base += 256 * 4;
    r.ring = reinterpret_cast<unsigned*>(base); base += 8 * r.pitch * 4;
    r.counters = reinterpret_cast<int*>(base);
    return r;
}

__device__ inline void rank_tables(const int* lv, RankWs w) {
    const int v = threadIdx.x;
    if (v < 256) {
        unsigned lo = 0, hi = 0;
        for (int k = 0; k < 3; ++k) lo |= (unsigned)(v <= lv[k]) << (10 * k);
        for (int k = 0; k < 3; ++k) hi |= (unsigned)(v <= lv[3 + k]) << (10 * k);
        w.lut0[v] = lo;
        w.lut1[v] = hi;
    }
    if (v < 4) w.counters[v] = 0;
}

// Exact decision for one pixel: #(window <= g+thr) <= 220 or #(window <= g-thr-1) >= 221.
__device__ inline bool rank_exact_pixel_thread(const uint8_t* gray, const Geom& g, int thr, int x, int y) {
    const int gv = gray[y * g.gp + x];
    const int pa = gv + thr, pb = gv - thr - 1;
    int ca = 0, cb = 0;
    for (int dy = -10; dy <= 10; ++dy) {
        const uint8_t* row = gray + min(max(y + dy, 0), g.h - 1) * g.gp;
        for (int dx = -10; dx <= 10; ++dx) {
            const int v = row[min(max(x + dx, 0), g.w - 1)];
            ca += v <= pa;
            cb += v <= pb;
        }
    }
    return ca <= 220 || cb >= 221;
}

// Per-pixel classification of one dirty cell (one thread): defect-for-sure pixels of
// the ROI go straight into CAND, ambiguous ones onto the exact list.
__device__ inline void rank_dirty_cell(const uint8_t* gray, const Geom& g, int thr, const unsigned* ROI, unsigned* CAND,
                                       RankWs& w, int ci, int cj, unsigned cw) {
    const unsigned u1 = cw & 255u, u2 = (cw >> 8) & 255u, u3 = (cw >> 16) & 255u, u4 = cw >> 24;
    for (int rr = 0; rr < kCell; ++rr) {
        const int y = cj * kCell + rr;
        if (y >= g.h) break;
        for (int cc = 0; cc < kCell; ++cc) {
            const int x = ci * kCell + cc;
            if (x >= g.w) break;
            const int wi = y * g.wpr + (x >> 5);
            const unsigned bit = 1u << (x & 31);
            if (!(ROI[wi] & bit)) continue;
            const unsigned gv = gray[y * g.gp + x];
            if (gv < u1 || gv > u4) {
                atomicOr(&CAND[wi], bit);
            } else if (!(gv >= u2 && gv <= u3)) {
                atomicAdd(&w.counters[2], 1);
                const int k = atomicAdd(&w.counters[1], 1);
                if (k < w.exact_cap) w.exact[k] = ((unsigned)y << 16) | (unsigned)x;
                else if (rank_exact_pixel_thread(gray, g, thr, x, y)) atomicOr(&CAND[wi], bit);
            }
        }
    }
}

// CAND (zeroed by the caller) receives every ROI pixel with |g - med| > thr.
// Returns the number of pixels that needed an exact rank count.
__device__ inline int rank_stage_lattice(CtaScratch& cs_, const uint8_t* gray, const Geom& g, RankWs w, const int* lv,
                                         int thr, const unsigned* ROI, unsigned* CAND, PhaseTimer& pt) {
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const int nly = (g.h + kCell - 1) / kCell, nlx = (g.w + kCell - 1) / kCell;
    const int ntrip = nlx + 6;
    const int wm1 = g.w - 1, hm1 = g.h - 1;
    const bool vact = tid < g.w;
    const int vx = vact ? tid : 0;
    const uint8_t* gcol = gray + vx;
    unsigned S0 = 0, S1 = 0;
    int b = -3;                                     // next 3-row block of this column
    // gray bytes of block b, loaded one block ahead so the table loads never wait on them
    unsigned q0 = gcol[0], q1 = q0, q2 = q0;
    for (int j0 = 0; j0 < nly; j0 += kLatBand) {
        const int j1 = min(j0 + kLatBand, nly);
        // ---- V: this column's blocks up to j1+2 ------------------------------------
        if (vact) {
            for (; b < j1 + 3; ++b) {
                const unsigned B0 = w.lut0[q0] + w.lut0[q1] + w.lut0[q2];
                const unsigned B1 = w.lut1[q0] + w.lut1[q1] + w.lut1[q2];
                const int rn = kCell * (b + 1);
                if (rn >= 0 && rn + 2 <= hm1) {                 // interior block: no clamping
                    const uint8_t* pr = gcol + rn * g.gp;
                    q0 = pr[0]; q1 = pr[g.gp]; q2 = pr[2 * g.gp];
                } else {
                    q0 = gcol[min(max(rn, 0), hm1) * g.gp];
                    q1 = gcol[min(max(rn + 1, 0), hm1) * g.gp];
                    q2 = gcol[min(max(rn + 2, 0), hm1) * g.gp];
                }
                const unsigned o = w.ring[((b + 1) & 7) * w.pitch + vx];
                S0 += B0; S1 += B1;
                if (b >= 4) { S0 -= o & kFld; S1 -= (o >> 2) & kFld; }
                w.ring[(b & 7) * w.pitch + vx] = B0 | (B1 << 2);
                if (b >= 3) w.cs[(b - 3 - j0) * w.pitch + vx] = make_uint2(S0, S1);
            }
        }
        __syncthreads();
        pt.acc(20);
        // ---- H: one warp per lattice row, one lane per cell --------------------------
        for (int jj = warp; jj < j1 - j0; jj += kWarps) {
            const uint2* row = w.cs + jj * w.pitch;
            const int cj = j0 + jj;
            // kPassBatch passes are independent: their loads and scans are interleaved for ILP
            for (int kb = 0; kb < nlx; kb += kCellsPerPass * kPassBatch) {
                unsigned p0[kPassBatch], p1[kPassBatch];
#pragma unroll
                for (int u = 0; u < kPassBatch; ++u) {
                    // triple k covers extended columns 3k+1..3k+3  <->  columns clamp(3k-9 .. 3k-7)
                    const int k = kb + u * kCellsPerPass + lane;
                    p0[u] = 0; p1[u] = 0;
                    if (k < ntrip && kb + u * kCellsPerPass < nlx) {
                        const int c0 = kCell * k - 9;
                        const uint2 a = row[min(max(c0, 0), wm1)], bb = row[min(max(c0 + 1, 0), wm1)],
                                    c = row[min(max(c0 + 2, 0), wm1)];
                        p0[u] = a.x + bb.x + c.x; p1[u] = a.y + bb.y + c.y;
                    }
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
                    for (int u = 0; u < kPassBatch; ++u) {
                        const unsigned x0 = __shfl_up_sync(kFull, p0[u], o), x1 = __shfl_up_sync(kFull, p1[u], o);
                        if (lane >= o) { p0[u] += x0; p1[u] += x1; }
                    }
                }
                unsigned C0[kPassBatch], C1[kPassBatch];
#pragma unroll
                for (int u = 0; u < kPassBatch; ++u) {
                    // window of cell i = k: triples i..i+6 = P[lane+6] - P[lane-1]
                    const unsigned h0 = __shfl_down_sync(kFull, p0[u], 6), h1 = __shfl_down_sync(kFull, p1[u], 6);
                    unsigned l0 = __shfl_up_sync(kFull, p0[u], 1), l1 = __shfl_up_sync(kFull, p1[u], 1);
                    if (lane == 0) { l0 = 0; l1 = 0; }
                    C0[u] = h0 - l0; C1[u] = h1 - l1;
                }
                bool dirty[kPassBatch];
                unsigned cw[kPassBatch];
                // the cell's 9 pixels (clamped at the crop edge: duplicates are harmless), all passes in flight
                int px[kPassBatch][9];
                const int y0 = kCell * cj;
                const uint8_t* r0 = gray + min(y0, hm1) * g.gp;
                const uint8_t* r1 = gray + min(y0 + 1, hm1) * g.gp;
                const uint8_t* r2 = gray + min(y0 + 2, hm1) * g.gp;
#pragma unroll
                for (int u = 0; u < kPassBatch; ++u) {
                    const int x0 = kCell * (kb + u * kCellsPerPass + lane);
                    const int xa = min(x0, wm1), xb = min(x0 + 1, wm1), xc = min(x0 + 2, wm1);
                    px[u][0] = r0[xa]; px[u][1] = r0[xb]; px[u][2] = r0[xc];
                    px[u][3] = r1[xa]; px[u][4] = r1[xb]; px[u][5] = r1[xc];
                    px[u][6] = r2[xa]; px[u][7] = r2[xb]; px[u][8] = r2[xc];
                }
#pragma unroll
                for (int u = 0; u < kPassBatch; ++u) {
                    const int k = kb + u * kCellsPerPass + lane;
                    const bool cact = lane < kCellsPerPass && k < nlx;
                    const int n263 = __popc((C0[u] + kGe263) & kFlag) + __popc((C1[u] + kGe263) & kFlag);
                    const int n179 = __popc((C0[u] + kGe179) & kFlag) + __popc((C1[u] + kGe179) & kFlag);
                    const int lo_idx = kLevels - n179;        // levels 0..lo_idx-1 surely have C <= 220: med >  LO
                    const int hi_idx = kLevels - n263;        // level hi_idx surely has C >= 221:        med <= HI
                    const int LO = lo_idx > 0 ? lv[lo_idx - 1] : -1;
                    const int HI = hi_idx < kLevels ? lv[hi_idx] : 255;
                    const int U1 = min(max(LO - thr + 1, 0), 255);     // defect if g <  U1
                    const int U4 = min(HI + thr, 255);                 // defect if g >  U4
                    const int U2 = max(HI - thr, 0);                   // clean needs g >= U2
                    const int U3 = min(LO + thr + 1, 255);             // clean needs g <= U3
                    cw[u] = (unsigned)U1 | ((unsigned)U2 << 8) | ((unsigned)U3 << 16) | ((unsigned)U4 << 24);
                    int mn = px[u][0], mx = px[u][0];
#pragma unroll
                    for (int t = 1; t < 9; ++t) { mn = min(mn, px[u][t]); mx = max(mx, px[u][t]); }
                    dirty[u] = cact && (mn < U2 || mx > U3);
                }
#pragma unroll
                for (int u = 0; u < kPassBatch; ++u) {
                    // append the dirty cells of this pass (one atomic per warp)
                    const unsigned dm = __ballot_sync(kFull, dirty[u]);
                    if (dm) {
                        const int k = kb + u * kCellsPerPass + lane;
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&w.counters[0], __popc(dm));
                        base = __shfl_sync(kFull, base, 0);
                        if (dirty[u]) {
                            const int slot = base + __popc(dm & ((1u << lane) - 1u));
                            if (slot < w.dirty_cap) w.dirty[slot] = make_uint2(((unsigned)cj << 16) | (unsigned)k, cw[u]);
                            else rank_dirty_cell(gray, g, thr, ROI, CAND, w, k, cj, cw[u]);
                        }
                    }
                }
            }
        }
        __syncthreads();
        pt.acc(21);
    }
    // ---- dirty cells: per-pixel classification ----------------------------------------
    const int nd = min(w.counters[0], w.dirty_cap);
    for (int i = tid; i < nd; i += kThreads) {
        const uint2 e = w.dirty[i];
        rank_dirty_cell(gray, g, thr, ROI, CAND, w, (int)(e.x & 0xffffu), (int)(e.x >> 16), e.y);
    }
    __syncthreads();
    pt.acc(22);
    // ---- ambiguous pixels: exact rank counts, one warp per pixel -------------------------
    const int ne = min(w.counters[1], w.exact_cap);
    for (int k2 = warp; k2 < ne; k2 += kWarps) {
        const unsigned ent = w.exact[k2];
        const int y = (int)(ent >> 16), x = (int)(ent & 0xffffu);
        const int gv = gray[y * g.gp + x];
        const int pa = gv + thr, pb = gv - thr - 1;
        int vals[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const int e = lane + 32 * k;
            const int dy = e / 21, dx = e - dy * 21;
            const int yy = min(max(y + dy - 10, 0), hm1), xx = min(max(x + dx - 10, 0), wm1);
            vals[k] = e < 441 ? (int)gray[yy * g.gp + xx] : 256;
        }
        unsigned ca = 0, cb = 0;
#pragma unroll
        for (int k = 0; k < 14; ++k) { ca += vals[k] <= pa; cb += vals[k] <= pb; }
        ca = __reduce_add_sync(kFull, ca);
        cb = __reduce_add_sync(kFull, cb);
        if (lane == 0 && (ca <= 220 || cb >= 221)) atomicOr(&CAND[y * g.wpr + (x >> 5)], 1u << (x & 31));
    }
    __syncthreads();
    const int total = w.counters[2];
    __syncthreads();
    (void)cs_;
    return total;
}

}  // namespace vi
