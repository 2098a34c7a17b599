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
// 3. Per pixel, four byte thresholds of its cell decide: defect for sure
//    (g <= LO-thr or g >= HI+thr+1), clean for sure (HI-thr <= g <= LO+thr+1), else
//    "ambiguous": those pixels (inside the ROI) get an exact rank count at their own
//    two pivots.  Exact for any level set and any image
//    (oracle/restate.py: residual_mask_lattice is the numpy twin).
//
// Box sums are separable.  V: one thread per column accumulates 3-row block sums
// (a lattice row's 21-row window is exactly 7 blocks), a ring of packed block sums
// gives the sliding 7-block sum.  H: one warp per lattice row turns the column
// sums (10 replicated columns each side) into a prefix in place; a cell's window
// sum is P[3i+21] - P[3i].
#pragma once
#include "vi_device.cuh"

namespace vi {

constexpr int kCell = 3;
constexpr int kLatBand = 16;             // lattice rows per band (= warps per CTA)
constexpr unsigned kFld = 0x00300C03u;   // 2-bit block-sum fields at the 10-bit field positions
constexpr unsigned kFlag = 0x20080200u;  // bit 9 of each 10-bit field
constexpr unsigned kGe263 = 249u | (249u << 10) | (249u << 20);   // field + 249 >= 512  <=>  field >= 263
constexpr unsigned kGe179 = 333u | (333u << 10) | (333u << 20);   // field + 333 >= 512  <=>  field >= 179

struct RankWs {
    uint2* lut;         // [256] packed indicators of a gray value
    unsigned* ring;     // [8][ring_pitch] packed 3-row block sums
    uint2* cs;          // [kLatBand][cs_pitch] column sums, then prefix sums over the replicate-extended row
    unsigned* cell;     // [kLatBand][cell_pitch] U1 | U2<<8 | U3<<16 | U4<<24
    int ring_pitch, cs_pitch, cell_pitch;
};

__host__ __device__ inline int rank_nlx(int w) { return (w + kCell - 1) / kCell; }
__host__ __device__ inline int rank_cs_pitch(int w) { return ((kCell * rank_nlx(w) + 22 + kSegL - 1) / kSegL) * kSegL; }
__host__ __device__ inline int rank_ws_bytes(int w) {
    return 256 * 8 + 8 * ((w + 3) & ~3) * 4 + kLatBand * rank_cs_pitch(w) * 8 + kLatBand * ((rank_nlx(w) + 3) & ~3) * 4 + 64;
}

__device__ inline RankWs rank_ws_carve(unsigned char* base, int w) {
    RankWs r;
    r.ring_pitch = (w + 3) & ~3;
    r.cs_pitch = rank_cs_pitch(w);
    r.cell_pitch = (rank_nlx(w) + 3) & ~3;
    r.lut = reinterpret_cast<uint2*>(base); base += 256 * 8;
    r.cs = reinterpret_cast<uint2*>(base); base += kLatBand * r.cs_pitch * 8;
    r.cell = reinterpret_cast<unsigned*>(base); base += kLatBand * r.cell_pitch * 4;
    r.ring = reinterpret_cast<unsigned*>(base);
    return r;
}

__device__ inline void rank_tables(const int* lv, RankWs w) {
    const int v = threadIdx.x;
    if (v < 256) {
        unsigned lo = 0, hi = 0;
        for (int k = 0; k < 3; ++k) lo |= (unsigned)(v <= lv[k]) << (10 * k);
        for (int k = 0; k < 3; ++k) hi |= (unsigned)(v <= lv[3 + k]) << (10 * k);
        w.lut[v] = make_uint2(lo, hi);
    }
}

// SURE / AMB: per-pixel bit masks for the whole unit (all pixels, ROI or not).
__device__ inline void rank_stage_lattice(const uint8_t* gray, const Geom& g, RankWs w, const int* lv, int thr,
                                          unsigned* SURE, unsigned* AMB, PhaseTimer& pt) {
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const int nly = (g.h + kCell - 1) / kCell, nlx = rank_nlx(g.w);
    const int nseg = w.cs_pitch / kSegL;            // <= 32
    const bool vact = tid < g.w;
    const int vx = vact ? tid : 0;
    const uint8_t* gcol = gray + vx;
    const int hm1 = g.h - 1;
    unsigned S0 = 0, S1 = 0;
    int b = -3;                                     // next 3-row block of this column
    // gray bytes of block b, loaded one block ahead so the LUT loads never wait on them
    unsigned q0 = gcol[0], q1 = q0, q2 = q0;
    for (int j0 = 0; j0 < nly; j0 += kLatBand) {
        const int j1 = min(j0 + kLatBand, nly);
        // ---- V: this column's blocks up to j1+2 ------------------------------------
        if (vact) {
            for (; b < j1 + 3; ++b) {
                const uint2 e0 = w.lut[q0], e1 = w.lut[q1], e2 = w.lut[q2];
                const int rn = kCell * (b + 1);
                if (rn >= 0 && rn + 2 <= hm1) {                 // interior block: no clamping
                    const uint8_t* pr = gcol + rn * g.gp;
                    q0 = pr[0]; q1 = pr[g.gp]; q2 = pr[2 * g.gp];
                } else {
                    q0 = gcol[min(max(rn, 0), hm1) * g.gp];
                    q1 = gcol[min(max(rn + 1, 0), hm1) * g.gp];
                    q2 = gcol[min(max(rn + 2, 0), hm1) * g.gp];
                }
                const unsigned o = w.ring[((b + 1) & 7) * w.ring_pitch + vx];
                const unsigned B0 = e0.x + e1.x + e2.x, B1 = e0.y + e1.y + e2.y;
                S0 += B0; S1 += B1;
                if (b >= 4) { S0 -= o & kFld; S1 -= (o >> 2) & kFld; }
                w.ring[(b & 7) * w.ring_pitch + vx] = B0 | (B1 << 2);
                if (b >= 3) w.cs[(b - 3 - j0) * w.cs_pitch + vx] = make_uint2(S0, S1);
            }
        }
        __syncthreads();
        pt.acc(20);
        // ---- H: one warp per lattice row: prefix over the replicate-extended row ---
        for (int jj = warp; jj < j1 - j0; jj += kWarps) {
            uint2* row = w.cs + jj * w.cs_pitch;
            const bool sact = lane < nseg;
            unsigned p0[kSegL], p1[kSegL];
            unsigned a0 = 0, a1 = 0;
#pragma unroll
            for (int t = 0; t < kSegL; ++t) {
                // extended index e = 11*lane + t  <->  column clamp(e - 10)
                int col = min(max(lane * kSegL + t - 10, 0), g.w - 1);
                uint2 v = sact ? row[col] : make_uint2(0u, 0u);
                a0 += v.x; a1 += v.y;
                p0[t] = a0; p1[t] = a1;
            }
            unsigned t0 = a0, t1 = a1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned x0 = __shfl_up_sync(kFull, t0, o);
                unsigned x1 = __shfl_up_sync(kFull, t1, o);
                if (lane >= o) { t0 += x0; t1 += x1; }
            }
            const unsigned off0 = t0 - a0, off1 = t1 - a1;
            __syncwarp();                            // every lane has read its columns before anyone overwrites
            if (sact) {
#pragma unroll
                for (int t = 0; t < kSegL; ++t) row[lane * kSegL + t] = make_uint2(p0[t] + off0, p1[t] + off1);
            }
            __syncwarp();
            for (int i = lane; i < nlx; i += 32) {
                uint2 hi = row[kCell * i + 21], lo = row[kCell * i];
                unsigned C0 = hi.x - lo.x, C1 = hi.y - lo.y;
                int n263 = __popc((C0 + kGe263) & kFlag) + __popc((C1 + kGe263) & kFlag);
                int n179 = __popc((C0 + kGe179) & kFlag) + __popc((C1 + kGe179) & kFlag);
                int lo_idx = kLevels - n179;        // levels 0..lo_idx-1 surely have C <= 220: med >  LO
                int hi_idx = kLevels - n263;        // level hi_idx surely has C >= 221:        med <= HI
                int LO = lo_idx > 0 ? lv[lo_idx - 1] : -1;
                int HI = hi_idx < kLevels ? lv[hi_idx] : 255;
                int U1 = min(max(LO - thr + 1, 0), 255);     // defect if g <  U1
                int U4 = min(HI + thr, 255);                 // defect if g >  U4
                int U2 = max(HI - thr, 0);                   // clean needs g >= U2
                int U3 = min(LO + thr + 1, 255);             // clean needs g <= U3
                w.cell[jj * w.cell_pitch + i] = (unsigned)U1 | ((unsigned)U2 << 8) | ((unsigned)U3 << 16) | ((unsigned)U4 << 24);
            }
        }
        __syncthreads();
        pt.acc(21);
        // ---- classify: one warp per lattice row = three pixel rows sharing the cell words
        for (int jj = warp; jj < j1 - j0; jj += kWarps) {
            const unsigned* crow = w.cell + jj * w.cell_pitch;
            const int yb = (j0 + jj) * kCell;
            for (int c0 = 0; c0 < g.wpr; c0 += 5) {
                unsigned gv[kCell][5], cE[5], cO[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int x = min((c0 + k) * 32 + lane, g.w - 1);
                    const unsigned cw = crow[x / kCell];
                    cE[k] = __byte_perm(cw, 0u, 0x4240);          // (U1, U3)
                    cO[k] = __byte_perm(cw, 0u, 0x4341);          // (U2, U4)
#pragma unroll
                    for (int rr = 0; rr < kCell; ++rr) gv[rr][k] = gray[min(yb + rr, hm1) * g.gp + x];
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int c = c0 + k;
                    if (c < g.wpr) {
                        const bool act = c * 32 + lane < g.w;
#pragma unroll
                        for (int rr = 0; rr < kCell; ++rr) {
                            if (yb + rr <= hm1) {
                                // per 16-bit field: (g + 0x200) - U has bit 9 set iff g >= U, (g + 0x1FF) - U iff g > U
                                const unsigned G2 = gv[rr][k] * 0x00010001u + 0x01FF0200u;
                                const unsigned dE = G2 - cE[k], dO = G2 - cO[k];
                                const bool ge1 = dE & 0x00000200u, gt3 = dE & 0x02000000u;
                                const bool ge2 = dO & 0x00000200u, gt4 = dO & 0x02000000u;
                                const bool df = !ge1 || gt4, ok = ge2 && !gt3;
                                const unsigned sure = __ballot_sync(kFull, act && df);
                                const unsigned amb = __ballot_sync(kFull, act && !df && !ok);
                                if (lane == 0) { SURE[(yb + rr) * g.wpr + c] = sure; AMB[(yb + rr) * g.wpr + c] = amb; }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        pt.acc(22);
    }
}

// Exact rank counts for the pixels set in Q: CAND |= pixel iff
// #(window <= g+thr) <= 220 or #(window <= g-thr-1) >= 221.  The pixels are
// compacted into `list` and each is evaluated by one warp (441 window pixels over
// 32 lanes).  Returns the number of pixels evaluated.
__device__ inline unsigned rank_exact_list(CtaScratch& cs, const uint8_t* gray, const Geom& g, int thr, unsigned* CAND,
                                           const unsigned* Q, unsigned* list, int cap) {
    const int per = (g.nwords + kThreads - 1) / kThreads;
    const int i0 = threadIdx.x * per, i1 = min(i0 + per, g.nwords);
    unsigned n = 0, dummy = 0, total, td;
    for (int i = i0; i < i1; ++i) n += __popc(Q[i]);
    unsigned off = n;
    cta_excl_scan2(cs, off, dummy, total, td);
    if (total == 0) return 0;
    const int lane = lane_id();
    for (unsigned base = 0; base < total; base += (unsigned)cap) {
        // (re)build the slice [base, base+cap) of the candidate list
        unsigned k = off;
        for (int i = i0; i < i1; ++i) {
            unsigned q = Q[i];
            int y = i / g.wpr, c = i - y * g.wpr;
            while (q) {
                int bpos = __ffs(q) - 1; q &= q - 1;
                if (k >= base && k < base + (unsigned)cap) list[k - base] = ((unsigned)y << 16) | (unsigned)(c * 32 + bpos);
                ++k;
            }
        }
        __syncthreads();
        const unsigned cnt = min((unsigned)cap, total - base);
        for (unsigned k2 = warp_id(); k2 < cnt; k2 += kWarps) {
            unsigned ent = list[k2];
            int y = (int)(ent >> 16), x = (int)(ent & 0xffffu);
            int gv = gray[y * g.gp + x];
            int pa = gv + thr, pb = gv - thr - 1;
            int vals[14];
#pragma unroll
            for (int k = 0; k < 14; ++k) {
                int e = lane + 32 * k;
                int dy = e / 21, dx = e - dy * 21;
                int yy = min(max(y + dy - 10, 0), g.h - 1), xx = min(max(x + dx - 10, 0), g.w - 1);
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
    }
    return total;
}

}  // namespace vi
