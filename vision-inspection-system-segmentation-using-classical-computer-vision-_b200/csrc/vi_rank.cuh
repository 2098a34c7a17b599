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
// Box sums are separable and every stage below is a straight-line pass:
//   V   one thread per column (whole cells per warp, spread over all warps) walks down the
//       rows in 3-row blocks: B(b) = sum of the packed level indicators of the
//       block (field-parallel arithmetic, no table), S(j) = sum of blocks j-3..j+3 (the
//       21-row window of lattice row j) kept by a ring of packed block sums, and
//       the block's gray min / max reduced over the cell's 3 columns by two
//       shuffles (min and 255-max packed as 16-bit halves: one VIMNMX3 does both).
//   S   one warp per lattice row turns S(j, x) into an inclusive prefix over x
//       (11 columns per lane, one warp scan; packed fields may wrap, differences
//       of prefixes are exact because every 21-column sum is < 1024).
//   C   one lane per cell: window = P[x+10] - P[x-11] (+ replicated edge columns),
//       two field-parallel compares give the bracket, a 49-entry table the four
//       byte thresholds, and the cell min / max decide clean or dirty.
#pragma once
#include "vi_pipeline.cuh"

namespace vi {

constexpr int kCell = 3;
constexpr int kLatBand = 16;             // lattice rows per band (= warps per CTA)
constexpr int kColsPerWarp = 30;         // V pass: 10 whole cells per warp
constexpr int kRankMaxW = kWarps * kColsPerWarp;      // widest unit of the lattice pass (480)
constexpr int kCmmRows = 2 * kLatBand;    // cell min/max rows: two band buffers (a band also writes the next band's first 3 rows)
constexpr unsigned kFld = 0x00300C03u;   // 2-bit block-sum fields at the 10-bit field positions
constexpr unsigned kFlag = 0x20080200u;  // bit 9 of each 10-bit field
constexpr unsigned kGe263 = 249u | (249u << 10) | (249u << 20);   // field + 249 >= 512  <=>  field >= 263
constexpr unsigned kGe179 = 333u | (333u << 10) | (333u << 20);   // field + 333 >= 512  <=>  field >= 179

struct RankWs {
    uint2* cs;          // [kLatBand][P] (0, then per column) 7-block column sums, then their prefix over x
    unsigned short* cmm;// [2][kLatBand][cpitch] cell gray min | (255 - max) << 8
    unsigned* table;    // [49] thresholds U1 | U2 << 8 | U3 << 16 | U4 << 24 by (n179, n263)
    uint2* dirty;       // [cells] (cell position, thresholds), in the CTA's global scratch: every cell fits
    unsigned* exact;    // [exact_cap] ambiguous pixels (y << 16 | x), global scratch; more are counted inline
    int exact_cap;
    int* counters;      // shared: [0] dirty cells, [1] ambiguous pixels listed, [2] ambiguous pixels total
    int P, cpitch;
    int dirty_cap;      // cells of the largest unit: what the dirty list holds
};

// Columns per lane of the prefix pass (odd: a lane's chunk starts land on distinct banks) and the row pitch of
// cs in uint2: one leading zero (the prefix "before column 0") + 32 * CH columns + one dummy column (written by
// the V lanes that own no column), so no access needs a bounds test (a prefix only flows forward: what lies
// past column w-1 is never read back).
__host__ __device__ inline int rank_ch(int w) { return w <= 352 ? 11 : 15; }
__host__ __device__ inline int rank_P(int w) { return 32 * rank_ch(w) + 3; }      // odd: see rank_group_cells
__host__ __device__ inline int rank_cpitch(int w) { return ((w + kCell - 1) / kCell + 2) & ~1; }      // cells + a dummy slot
__host__ __device__ inline int rank_ws_bytes(int w) {
    return kLatBand * rank_P(w) * 8 + kCmmRows * rank_cpitch(w) * 2 + 64 * 4;
}

constexpr int kExactCap = 4096;
constexpr int kExactWarpMax = 192;       // more ambiguous pixels than this: one thread per pixel instead of one warp
__host__ __device__ inline long long rank_scratch_bytes(int wmax, int hmax) {
    return (long long)((wmax + kCell - 1) / kCell) * ((hmax + kCell - 1) / kCell) * 8 + kExactCap * 4;
}

// The cell pass runs before the ROI exists (next to the Otsu scan), the per-pixel pass after it: the two lists
// outlive the shared workspace, so they sit in the CTA's global scratch (L2-resident) and the counters in static
// shared memory.
__device__ inline RankWs rank_ws_carve(unsigned char* base, int w, unsigned char* lists, int wmax, int hmax, int* counters) {
    RankWs r;
    r.P = rank_P(w); r.cpitch = rank_cpitch(w);
    r.cs = reinterpret_cast<uint2*>(base); base += kLatBand * r.P * 8;
    r.cmm = reinterpret_cast<unsigned short*>(base); base += kCmmRows * r.cpitch * 2;
    r.table = reinterpret_cast<unsigned*>(base);
    r.counters = counters;
    r.dirty = reinterpret_cast<uint2*>(lists);
    r.exact = reinterpret_cast<unsigned*>(lists + (long long)((wmax + kCell - 1) / kCell) * ((hmax + kCell - 1) / kCell) * 8);
    r.exact_cap = kExactCap;
    r.dirty_cap = ((wmax + kCell - 1) / kCell) * ((hmax + kCell - 1) / kCell);
    return r;
}

__device__ inline void rank_dirty_cell(const uint8_t* gray, const Geom& g, int thr, const unsigned* ROI, unsigned* CAND,
                                       RankWs& w, int ci, int cj, unsigned cw);

// Warp 0 only (the caller synchronises): the threshold table by bracket, the list counters, the zero column of cs.
__device__ inline void rank_tables(const int* lv, int thr, RankWs w) {
    const int lane = lane_id();
    for (int e = lane; e < 49; e += 32) {
        const int n179 = e / 7, n263 = e - n179 * 7;
        const int lo_idx = kLevels - n179;        // levels 0..lo_idx-1 surely have C <= 220: med >  LO
        const int hi_idx = kLevels - n263;        // level hi_idx surely has C >= 221:        med <= HI
        const int LO = lo_idx > 0 ? lv[lo_idx - 1] : -1;
        const int HI = hi_idx < kLevels ? lv[hi_idx] : 255;
        const int U1 = min(max(LO - thr + 1, 0), 255);     // defect if g <  U1
        const int U4 = min(HI + thr, 255);                 // defect if g >  U4
        const int U2 = max(HI - thr, 0);                   // clean needs g >= U2
        const int U3 = min(LO + thr + 1, 255);             // clean needs g <= U3
        w.table[e] = (unsigned)U1 | ((unsigned)U2 << 8) | ((unsigned)U3 << 16) | ((unsigned)U4 << 24);
    }
    if (lane < 4) w.counters[lane] = 0;
    if (lane < kLatBand) w.cs[lane * w.P] = make_uint2(0u, 0u);      // the prefix before column 0
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
    // most dirty cells (plate edges, bright field) lie outside the ROI: three ROI bits per row decide that first
    const int x0 = ci * kCell, c0 = x0 >> 5, sh = x0 & 31;
    unsigned roi3[kCell];
    unsigned any = 0;
#pragma unroll
    for (int rr = 0; rr < kCell; ++rr) {
        const int y = min(cj * kCell + rr, g.h - 1);
        const unsigned* row = ROI + y * g.wpr;
        unsigned b = row[c0] >> sh;
        if (sh > 29 && c0 + 1 < g.wpr) b |= row[c0 + 1] << (32 - sh);
        roi3[rr] = (cj * kCell + rr < g.h) ? (b & 7u) : 0u;          // bits past the crop edge are 0 in every mask
        any |= roi3[rr];
    }
    if (!any) return;
#pragma unroll
    for (int rr = 0; rr < kCell; ++rr) {
        const int y = cj * kCell + rr;
        if (y >= g.h) break;
#pragma unroll
        for (int cc = 0; cc < kCell; ++cc) {
            const int x = ci * kCell + cc;
            if (x >= g.w) break;
            const int wi = y * g.wpr + (x >> 5);
            const unsigned bit = 1u << (x & 31);
            if (!((roi3[rr] >> cc) & 1u)) continue;
            const unsigned gv = gray[y * g.gp + x];
            if (gv < u1 || gv > u4) {
                atomicOr(&CAND[wi], bit);
            } else if (!(gv >= u2 && gv <= u3)) {
                atomicAdd(&w.counters[2], 1);
                const int k = atomicAdd(&w.counters[1], 1);
                VI_CHECK(k >= 0 && y < g.h && x < g.w, CHK_EXACT_LIST);
                if (k < w.exact_cap) w.exact[k] = ((unsigned)y << 16) | (unsigned)x;
                else if (rank_exact_pixel_thread(gray, g, thr, x, y)) atomicOr(&CAND[wi], bit);
            }
        }
    }
}

// gray q as (q, 255 - q) 16-bit halves: the packed minimum carries min and 255 - max.
__device__ __forceinline__ unsigned mm_pack(unsigned q) { return q * 0xFFFF0001u + 0x00FF0000u; }

// S + C for a group of kGrp consecutive cells of one lattice row (one thread).  The window of cell i is columns
// 3i-9 .. 3i+11 = the seven 3-column cells i-3 .. i+3 (columns outside the crop replicate the edge column: clamped
// indices), so a thread adds 3 * (kGrp + 6) column sums into cell sums, forms the first window from seven of them and
// slides it over the group.  No prefix pass, no cross-lane traffic: every load is independent of the others.
// Threads of a half-warp hold the 16 rows of a band (the row pitch P is odd: their 8-byte loads hit distinct banks).
constexpr int kGrp = 4;
static_assert(kLatBand == 16, "the S+C pass maps the 16 lanes of a half-warp to the rows of a band");

template <bool EDGE>
__device__ __forceinline__ void rank_group_cells(const Geom& g, RankWs& w, const uint2* row, const unsigned short* cm, int cj,
                                                 int gi, int nlx, bool active) {
    const int wm1 = g.w - 1;
    const int c0 = kCell * kGrp * gi - 9;
    uint2 cs[kGrp + 6];
#pragma unroll
    for (int k = 0; k < kGrp + 6; ++k) {
        const int c = c0 + kCell * k;
        uint2 t0, t1, t2;
        if (EDGE) { t0 = row[min(max(c, 0), wm1)]; t1 = row[min(max(c + 1, 0), wm1)]; t2 = row[min(max(c + 2, 0), wm1)]; }
        else { t0 = row[c]; t1 = row[c + 1]; t2 = row[c + 2]; }
        cs[k] = make_uint2(t0.x + t1.x + t2.x, t0.y + t1.y + t2.y);
    }
    unsigned C0 = 0, C1 = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) { C0 += cs[k].x; C1 += cs[k].y; }
    const int lane = lane_id();
#pragma unroll
    for (int m = 0; m < kGrp; ++m) {
        const int i = kGrp * gi + m;
        const int ic = min(i, nlx - 1);
        const int n263 = __popc((C0 + kGe263) & kFlag) + __popc((C1 + kGe263) & kFlag);
        const int n179 = __popc((C0 + kGe179) & kFlag) + __popc((C1 + kGe179) & kFlag);
        const unsigned cw = w.table[n179 * 7 + n263];
        const unsigned mmv = cm[ic];
        const unsigned mn = mmv & 255u, mx = 255u - (mmv >> 8);
        const bool dirty = active && i < nlx && (mn < ((cw >> 8) & 255u) || mx > ((cw >> 16) & 255u));
        const unsigned dm = __ballot_sync(kFull, dirty);
        if (dm) {                                                    // append the dirty cells (one atomic per warp)
            int base = 0;
            if (lane == 0) base = atomicAdd(&w.counters[0], __popc(dm));
            base = __shfl_sync(kFull, base, 0);
            VI_CHECK(base >= 0 && base + __popc(dm) <= w.dirty_cap, CHK_DIRTY_LIST);
            if (dirty) w.dirty[base + __popc(dm & ((1u << lane) - 1u))] = make_uint2(((unsigned)cj << 16) | (unsigned)i, cw);
        }
        if (m + 1 < kGrp) { C0 += cs[m + 7].x - cs[m].x; C1 += cs[m + 7].y - cs[m].y; }
    }
}

// One 3-row block of one column: packed level counts of the block and its gray (min, 255 - max).
struct VBlk { unsigned B0, B1, mm; };

// Level indicators without a table (table loads with data-dependent addresses were bank-conflict bound):
// with R = 1 | 1<<10 | 1<<20 and T = sum_k (lv_k + 512) << 10k, field k of T - q*R is lv_k + 512 - q in
// [257, 766], so its bit 9 is (q <= lv_k).  The three pixels' bits are added by a full adder on whole words.
constexpr unsigned kRep = 0x00100401u;
constexpr unsigned kBit9 = 0x20080200u;

__device__ __forceinline__ unsigned ind3(unsigned T, unsigned x0, unsigned x1, unsigned x2) {     // x = q * kRep
    const unsigned a = T - x0, b = T - x1, c = T - x2;
    const unsigned lo = (a ^ b ^ c) & kBit9, hi = ((a & b) | (c & (a | b))) & kBit9;
    return (lo + 2 * hi) >> 9;
}

__device__ __forceinline__ VBlk v_block(unsigned T0, unsigned T1, unsigned q0, unsigned q1, unsigned q2) {
    VBlk r;
    const unsigned x0 = q0 * kRep, x1 = q1 * kRep, x2 = q2 * kRep;
    r.B0 = ind3(T0, x0, x1, x2);
    r.B1 = ind3(T1, x0, x1, x2);
    const unsigned mn = __vimin3_u32(q0, q1, q2), mx = __vimax3_u32(q0, q1, q2);
    r.mm = mn + ((255u - mx) << 16);
    return r;
}

// Min / max of a cell's 3 columns (lanes 3c, 3c+1, 3c+2 of the warp), as min | (255 - max) << 8.
__device__ __forceinline__ unsigned short cell_mm(unsigned mm) {
    const unsigned m1 = __shfl_down_sync(kFull, mm, 1), m2 = __shfl_down_sync(kFull, mm, 2);
    mm = __vimin3_u16x2(mm, m1, m2);
    return (unsigned short)((mm & 0xFFu) | (mm >> 8));
}

// Column state of the V pass: the sliding 7-block sums and the ring of the last 8 block sums
// (registers: the band loop is unrolled so every ring index is static).
struct VState {
    unsigned S0, S1;
    unsigned r0[8], r1[8];
    unsigned q0, q1, q2;       // gray of the next block, loaded one block ahead
};

// Blocks of one band: lattice rows j0 .. j0+15 <-> blocks b = j0+3 .. j0+18 (sequence n = b+3, ring slot n & 7).
template <bool CLAMP>
__device__ __forceinline__ void v_band(VState& st, unsigned T0, unsigned T1, const uint8_t* gcol, int gp, int hm1, int j0,
                                       uint2* csp, int P, unsigned short* cmA, unsigned short* cmB, int cpitch) {
    // every stride is an opaque register (the compiler otherwise re-derives them from the unit width per store)
    const uint8_t* pr = gcol + (kCell * (j0 + 4)) * gp;            // rows of block j0+4 (the first prefetch)
    int rn = kCell * (j0 + 4);
    unsigned short* cmp = cmA + 3 * cpitch;
#pragma unroll
    for (int u = 0; u < kLatBand; ++u) {
        const VBlk k = v_block(T0, T1, st.q0, st.q1, st.q2);        // block b = j0+3+u
        if (CLAMP) {
            st.q0 = gcol[min(rn, hm1) * gp]; st.q1 = gcol[min(rn + 1, hm1) * gp]; st.q2 = gcol[min(rn + 2, hm1) * gp];
            rn += kCell;
        } else {
            st.q0 = pr[0]; st.q1 = pr[gp]; st.q2 = pr[2 * gp];
            pr += kCell * gp;
        }
        *cmp = cell_mm(k.mm);                                       // lanes that lead no cell write a dummy slot
        cmp += cpitch;
        if (u == 12) cmp = cmB;
        const int slot = (6 + u) & 7, old = (7 + u) & 7;
        st.S0 += k.B0 - st.r0[old]; st.S1 += k.B1 - st.r1[old];
        st.r0[slot] = k.B0; st.r1[slot] = k.B1;
        *csp = make_uint2(st.S0, st.S1);                            // lanes without a column write the dummy column
        csp += P;
    }
}

// Part 1 (needs only the gray crop and the levels): window counts on the lattice, list of dirty cells.
// The warp kOtsuWarp carries the exact Otsu scan along when it owns no column: q1 sums and reciprocals during the
// first band, a slice of the mu1 recurrence during each further band, the rest at the end; *otsu_t holds the
// threshold on return.
template <class PT>
VI_PHASE void rank_cells(const uint8_t* gray, const Geom& g, RankWs w, const int* lv, const unsigned* hist, int npix,
                         double* ows, int olast, int* otsu_t, PT& pt) {
    const int lane = lane_id(), warp = warp_id();
    const int nly = (g.h + kCell - 1) / kCell, nlx = (g.w + kCell - 1) / kCell;
    const int hm1 = g.h - 1;
    // V: the first 3*cpw lanes of warp v own columns 3*cpw*v .. (cpw whole cells); columns past the
    // crop edge re-read the edge column (duplicates never change a cell's min / max)
    // cells per warp: full warps for wide units (the pass is issue bound there); narrow units spread their few
    // columns over all warps instead (latency bound: 96 columns on 4 warps left 12 idle)
    const int cpw = nlx > 64 ? kColsPerWarp / kCell : max((nlx + kWarps - 1) / kWarps, 1);
    const int vcol = warp * cpw * kCell + lane;
    const bool vact = lane < cpw * kCell && vcol < g.w;
    const bool vwarp = warp * cpw * kCell < g.w;               // warps without a column skip the pass
    const uint8_t* gcol = gray + min(vcol, g.w - 1);
    const int vcell = warp * cpw + lane / kCell;
    const bool cell_lead = vact && (lane % kCell) == 0;
    const int bufrows = kLatBand * w.cpitch;
    VState st;
    const unsigned T0 = (unsigned)(lv[0] + 512) | ((unsigned)(lv[1] + 512) << 10) | ((unsigned)(lv[2] + 512) << 20);
    const unsigned T1 = (unsigned)(lv[3] + 512) | ((unsigned)(lv[4] + 512) << 10) | ((unsigned)(lv[5] + 512) << 20);
    if (vwarp) {
        // prologue: blocks b = -3 .. 2 (rows above the crop replicate row 0)
        const int gp = g.gp;
        const unsigned g0 = gcol[0];
        const VBlk ka = v_block(T0, T1, g0, g0, g0);
        st.S0 = 3 * ka.B0; st.S1 = 3 * ka.B1;
        st.r0[0] = st.r0[1] = st.r0[2] = ka.B0; st.r1[0] = st.r1[1] = st.r1[2] = ka.B1;
        st.r0[6] = st.r0[7] = 0; st.r1[6] = st.r1[7] = 0;
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const VBlk k = v_block(T0, T1, gcol[min(3 * bb, hm1) * gp], gcol[min(3 * bb + 1, hm1) * gp], gcol[min(3 * bb + 2, hm1) * gp]);
            st.S0 += k.B0; st.S1 += k.B1;
            st.r0[3 + bb] = k.B0; st.r1[3 + bb] = k.B1;
            const unsigned short cm = cell_mm(k.mm);
            w.cmm[bb * w.cpitch + (cell_lead ? vcell : w.cpitch - 1)] = cm;
        }
        st.q0 = gcol[min(9, hm1) * gp]; st.q1 = gcol[min(10, hm1) * gp]; st.q2 = gcol[min(11, hm1) * gp];
    }
    int P = w.P, cpitch = w.cpitch, gp = g.gp;
    asm volatile("" : "+r"(P), "+r"(cpitch), "+r"(gp));
    const int csx = vact ? 1 + vcol : P - 1;                  // column slot (dummy for lanes without a column)
    const int cmx = cell_lead ? vcell : cpitch - 1;           // cell slot (dummy for lanes that lead no cell)
    VI_CHECK(csx >= 1 && csx < P && cmx >= 0 && cmx < cpitch && 1 + g.w <= P - 1 && nlx <= cpitch - 1, CHK_LATTICE_SLOT);
    OtsuJob job;
    const bool owarp = warp == kOtsuWarp;
    const bool oslice = owarp && !vwarp;                      // idle during V: the scan rides along
    const int nbands = (nly + kLatBand - 1) / kLatBand;
    int per = 256;
    int par = 0;
    for (int j0 = 0; j0 < nly; j0 += kLatBand, par ^= 1) {
        const int j1 = min(j0 + kLatBand, nly);
        if (oslice) {
            if (j0 == 0) { otsu_begin(job, hist, npix, ows, olast); per = nbands > 1 ? (job.imax - job.imin + nbands - 1) / (nbands - 1) : 256; }
            else otsu_chain(job, per);
        }
        // ---- V: blocks j0+3 .. j0+18 of every column -----------------------------------
        if (vwarp) {
            unsigned short* cmA = w.cmm + par * bufrows + cmx;
            unsigned short* cmB = w.cmm + (par ^ 1) * bufrows + cmx;
            if (kCell * (j0 + kLatBand + 3) + 2 <= hm1)
                v_band<false>(st, T0, T1, gcol, gp, hm1, j0, w.cs + csx, P, cmA, cmB, cpitch);
            else
                v_band<true>(st, T0, T1, gcol, gp, hm1, j0, w.cs + csx, P, cmA, cmB, cpitch);
        }
        cta_sync();
        pt.acc(20);
        // ---- S + C: one thread per (lattice row of the band, group of kGrp cells) --------------
        {
            const int ngrp = (nlx + kGrp - 1) / kGrp;
            const int jj = threadIdx.x & (kLatBand - 1);
            const bool rowok = jj < j1 - j0;
            const uint2* row = w.cs + (rowok ? jj : 0) * w.P + 1;
            const unsigned short* cm = w.cmm + par * bufrows + (rowok ? jj : 0) * w.cpitch;
            for (int gb = 0; gb < ngrp; gb += kThreads / kLatBand) {
                const int gi = gb + (threadIdx.x >> 4);                     // uniform over a half-warp
                const bool active = rowok && gi < ngrp;
                const int gic = min(gi, ngrp - 1);
                const int c0 = kCell * kGrp * gic - 9;
                // warp-uniform choice (the two half-warps hold different groups; the ballots inside need the whole warp)
                const bool inner = __all_sync(kFull, c0 >= 0 && c0 + kCell * (kGrp + 6) - 1 <= g.w - 1);
                if (inner) rank_group_cells<false>(g, w, row, cm, j0 + jj, gic, nlx, active);
                else rank_group_cells<true>(g, w, row, cm, j0 + jj, gic, nlx, active);
            }
        }
        cta_sync();
        pt.acc(21);
    }
    if (owarp) {
        if (!oslice) otsu_begin(job, hist, npix, ows, olast);
        const int t = otsu_end(job);
        if (lane == 0) *otsu_t = t;
    }
    cta_sync();
}

// Part 2 (needs the ROI): CAND (zeroed by the caller) receives every ROI pixel with |g - med| > thr.
// Returns the number of pixels that needed an exact rank count.
template <class PT>
VI_PHASE int rank_finish(const uint8_t* gray, const Geom& g, RankWs w, int thr, const unsigned* ROI, unsigned* CAND,
                         PT& pt) {
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const int hm1 = g.h - 1, wm1 = g.w - 1;
    // ---- dirty cells: per-pixel classification ----------------------------------------
    const int nd = w.counters[0];
    VI_CHECK(nd >= 0 && nd <= w.dirty_cap, CHK_DIRTY_LIST);
    for (int i = tid; i < nd; i += kThreads) {
        const uint2 e = w.dirty[i];
        rank_dirty_cell(gray, g, thr, ROI, CAND, w, (int)(e.x & 0xffffu), (int)(e.x >> 16), e.y);
    }
    cta_sync();
    pt.acc(22);
    // ---- ambiguous pixels: exact rank counts ------------------------------------------------
    // Few of them (the usual case): one warp per pixel, 14 window pixels per lane.  Many (low thresholds, small
    // units whose windows mostly straddle the plate edge): one thread per pixel walks its own window -- a quarter
    // of the warp-instructions per pixel once the warps are full; the list is in cell order, so the lanes of a
    // warp read neighbouring windows.
    const int ne = min(w.counters[1], w.exact_cap);
    if (ne > kExactWarpMax) {
        const int gp = g.gp;
        for (int k2 = tid; k2 < ne; k2 += kThreads) {
            const unsigned ent = w.exact[k2];
            const int y = (int)(ent >> 16), x = (int)(ent & 0xffffu);
            const int gv = gray[y * gp + x];
            const int pa = gv + thr, pb = gv - thr - 1;
            int ca = 0, cb = 0;
            if (x >= 10 && x + 10 <= wm1 && y >= 10 && y + 10 <= hm1) {
                // window inside the crop: six aligned words per row (the crop pitch is a multiple of 4, so the byte
                // phase o is the same on every row), bytes compared four at a time; the per-byte "greater" flags are
                // summed as byte counters (at most 126 per lane) and the counts are 441 minus their totals.
                const int a0 = (y - 10) * gp + (x - 10);
                const int o = a0 & 3;
                const unsigned* wp = reinterpret_cast<const unsigned*>(gray + (a0 - o));
                const unsigned fm = 0x01010101u << (8 * o), lm = 0x01010101u >> (8 * (3 - o));      // first / last word bytes
                const SwarPivot qa = swar_pivot(pa), qb = swar_pivot(pb);
                unsigned ga = 0, gb = 0;
                const int wpitch = gp >> 2;
                for (int dy = 0; dy < 21; ++dy) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const unsigned W = wp[k];
                        const unsigned m = k == 0 ? fm : (k == 5 ? lm : 0x01010101u);
                        ga += swar_gt(W, qa) & m;
                        gb += swar_gt(W, qb) & m;
                    }
                    wp += wpitch;
                }
                const unsigned sa = (ga & 0x00ff00ffu) + ((ga >> 8) & 0x00ff00ffu), sb = (gb & 0x00ff00ffu) + ((gb >> 8) & 0x00ff00ffu);
                ca = 441 - (int)((sa & 0xffffu) + (sa >> 16));         // (the four lanes can add up to 441: no byte-wide total)
                cb = 441 - (int)((sb & 0xffffu) + (sb >> 16));
            } else {
                // window clipped to the crop; the replicated border rows / columns enter as weights of the edge ones
                const int xa = max(x - 10, 0), xb = min(x + 10, wm1), ya = max(y - 10, 0), yb = min(y + 10, hm1);
                const int nl = xa - (x - 10), nr = (x + 10) - xb, mt = ya - (y - 10), mb = (y + 10) - yb;
                for (int r = ya; r <= yb; ++r) {
                    const uint8_t* row = gray + r * gp;
                    const int e0 = row[0], e1 = row[wm1];
                    int ra = nl * (e0 <= pa) + nr * (e1 <= pa), rb = nl * (e0 <= pb) + nr * (e1 <= pb);
                    for (int c = xa; c <= xb; ++c) { const int v = row[c]; ra += v <= pa; rb += v <= pb; }
                    const int wr = 1 + (r == 0 ? mt : 0) + (r == hm1 ? mb : 0);
                    ca += wr * ra; cb += wr * rb;
                }
            }
            if (ca <= 220 || cb >= 221) atomicOr(&CAND[y * g.wpr + (x >> 5)], 1u << (x & 31));
        }
    } else
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
    cta_sync();
    const int total = w.counters[2];
    cta_sync();
    return total;
}

}  // namespace vi
