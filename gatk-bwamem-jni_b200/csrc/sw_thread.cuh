// sw_thread.cuh -- mate-rescue local alignment (ksw_align2 -> ksw_u8 of un-vendored lh3/bwa; reached
// from mem_matesw behind /root/reference/src/main/c/jnibwa.c:214) as INTER-TASK Smith-Waterman: one thread per task,
// two cells per 32-bit register, the row recurrence in Blackwell DPX instructions.
//
// The warp-per-task kernel (rescue.cuh warp_sw_fast) keeps a row in the registers of 32 lanes, five cells per lane: the
// fixed cost of a row (two lane scans, shuffles, reductions) is spread over five cells and the kernel spends ~80
// instructions per cell.  Here a thread owns a whole task in the striped layout of the SSE2 kernel it reproduces:
//   * a row is slen "vectors" of 16 cells (cell (s, j) = query column s * slen + j); a vector is eight 32-bit words of
//     two signed 16-bit halves, kept in shared memory interleaved over the threads of the block (bank tid % 32);
//   * per word and row:  h  = max3(min(diag + (score + shift), 255) - shift, e, f)      VIADDMNMX.S16x2, VIADD, VIMNMX3.S16x2
//                        e' = max(h - oe_del, e - e_del, 0)                             VIADD, VIADDMNMX.S16x2.RELU
//                        f' = max(h - oe_ins, f - e_ins, 0)                             VIADD, VIADDMNMX.S16x2.RELU
//     (the byte kernel's saturating arithmetic restated in 16 bits; scores of two cells come from one PRMT over the five
//     scores of the row's target base, selected by the two query codes);
//   * F runs along j inside a segment and restarts at 0 at every segment -- exactly the striped kernel's first pass; the
//     F leaving a segment is then carried into the following ones (its "lazy F" loop: oracle sw_striped, pass 2), which
//     is skipped on the rows where no segment has one.
// ~8 instructions per cell, and the 32 lanes of a warp run 32 tasks of (nearly) the same shape in lock-step.
// Tasks this formulation does not take (16-bit kernel: l_ms * a >= 250; queries beyond the shared rows) stay with the
// warp-per-task kernel.  Results are identical by construction: same recurrence, same order of rows.
#pragma once
#include "rescue.cuh"

struct TswMem {
	u32* H; u32* E; u16* sel; int stride;  // element k of this thread at [k * stride]
	int capw;                              // words per array (>= slen * 8)
	u32* hmax; int hmax_stride;            // snapshot of the best row: global, element k at [k * hmax_stride]
	u64* brow; int brow_cap;               // (row maximum, row) list: global, per thread slot
};

DEV u32 pk2(int v) { return (u32)(u16)(i16)v | (u32)(u16)(i16)v << 16; }
DEV int half_lo(u32 v) { return (int)(i16)(u16)(v & 0xffffu); }
DEV int half_hi(u32 v) { return (int)(i16)(u16)(v >> 16); }

// One pass of the byte kernel (forward: the whole query; reverse: the reversed query prefix against the reversed target
// prefix with KSW_XSTOP).  swtab[t]: byte c = score of query code c against target base t, plus shift (code 5 = padding
// column: shift alone); bytes 6, 7 = 0.
DEV SwFwd tsw_pass_u8(const DevIndex& x, const MemOpt& opt, const TswMem& tm, const u64* swtab, int shift, int mmax,
                      int qlen, const QAcc& qa, int tlen, const TAcc& ta, int xtra, u64& cells)
{
	SwFwd r; r.score = 0; r.te = -1; r.qe = -1; r.score2 = -1; r.te2 = -1;
	const int slen = (qlen + 15) >> 4, nw = slen << 3;
	if (slen == 0) return r;
	constexpr size_t ST = TX_STRIDE; // (== tm.stride)
	const int e_ins = opt.e_ins;
	const u32 NEG_EDEL = pk2(-opt.e_del), NEG_OEDEL = pk2(-(opt.o_del + opt.e_del)), NEG_EINS = pk2(-opt.e_ins), NEG_OEINS = pk2(-(opt.o_ins + opt.e_ins));
	const u32 NEG_SHIFT = pk2(-shift), SAT = 0x00ff00ffu;
	const int minsc = (xtra & KSW_XSUBO) ? (xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & KSW_XSTOP) ? (xtra & 0xffff) : 0x10000;
	for (int k = 0; k < nw; ++k) { tm.H[(size_t)k * ST] = 0u; tm.E[(size_t)k * ST] = 0u; }
	for (int j = 0; j < slen; ++j) {
#pragma unroll
		for (int w = 0; w < 8; ++w) {
			const int q0 = (2 * w) * slen + j, q1 = (2 * w + 1) * slen + j;
			const u32 c0 = q0 < qlen ? (u32)qacc_code(qa, q0) : 5u, c1 = q1 < qlen ? (u32)qacc_code(qa, q1) : 5u;
			tm.sel[(size_t)(j * 8 + w) * ST] = (u16)(c0 | 7u << 4 | c1 << 8 | 7u << 12); // bytes: score(c0), 0, score(c1), 0
		}
	}
	const i64 l_pac = x.l_pac;
	int gmax = 0, te = -1, n_b = 0;
	u32 pac_word = 0; i64 pac_widx = -1;
	for (int i = 0; i < tlen; ++i) {
		const i64 p = tacc_pos(ta, i);
		const bool rv = p >= l_pac;
		const i64 fp = rv ? (l_pac << 1) - 1 - p : p;
		const i64 widx = fp >> 4;
		if (widx != pac_widx) { pac_word = ldg((const u32*)x.pac + widx); pac_widx = widx; }
		int tb = (int)((pac_word >> (8 * (int)((fp >> 2) & 3))) >> ((~fp & 3) << 1) & 3);
		if (rv) tb = 3 - tb;
		const u64 tab = swtab[tb];
		const u32 tab_lo = (u32)tab, tab_hi = (u32)(tab >> 32);
		cells += (u64)qlen;
		u32 f[8], rmax[8], hd[8];
		{ // diagonal of column 0 of every segment: the last column of the segment before it (0 for the first)
			u32 prev = 0u;
#pragma unroll
			for (int w = 0; w < 8; ++w) { const u32 cur = tm.H[(size_t)((slen - 1) * 8 + w) * ST]; hd[w] = byte_perm(prev, cur, 0x5432u); prev = cur; f[w] = 0u; rmax[w] = 0u; }
		}
		for (int j = 0; j < slen; ++j) {
#pragma unroll
			for (int w = 0; w < 8; ++w) {
				const size_t k = (size_t)(j * 8 + w) * ST;
				const u32 hold = tm.H[k], e = tm.E[k];
				const u32 sc = byte_perm(tab_lo, tab_hi, (u32)tm.sel[k]);
				const u32 t = v_add2(dpx_addmin_s16x2(hd[w], sc, SAT), NEG_SHIFT);
				const u32 h = dpx_max3_s16x2(t, e, f[w]); // (e, f >= 0: the floor at 0 is implied)
				rmax[w] = v_maxs2(rmax[w], h);
				tm.H[k] = h;
				tm.E[k] = dpx_addmax_relu_s16x2(h, NEG_OEDEL, v_add2(e, NEG_EDEL));
				f[w] = dpx_addmax_relu_s16x2(h, NEG_OEINS, v_add2(f[w], NEG_EINS));
				hd[w] = hold;
			}
		}
		int imax;
		{
			u32 m = rmax[0];
#pragma unroll
			for (int w = 1; w < 8; ++w) m = v_maxs2(m, rmax[w]);
			imax = half_lo(m) > half_hi(m) ? half_lo(m) : half_hi(m);
		}
		// second pass: the F leaving a segment carried into the following segments
		if ((f[0] | f[1] | f[2] | f[3] | f[4] | f[5] | f[6] | f[7]) != 0u) {
			u32 C[8];
			int c = 0;
#pragma unroll
			for (int s = 0; s < 16; ++s) {
				if (s & 1) C[s >> 1] |= (u32)(u16)c << 16; else C[s >> 1] = (u32)(u16)c;
				const int fe = (s & 1) ? half_hi(f[s >> 1]) : half_lo(f[s >> 1]);
				c -= slen * e_ins; c = c > 0 ? c : 0; c = c > fe ? c : fe;
			}
			for (int j = 0; j < slen; ++j) {
				u32 any = 0u;
#pragma unroll
				for (int w = 0; w < 8; ++w) {
					const size_t k = (size_t)(j * 8 + w) * ST;
					tm.H[k] = v_maxs2(tm.H[k], C[w]);
					C[w] = dpx_addmax_relu_s16x2(C[w], NEG_EINS, 0u);
					any |= C[w];
				}
				if (!any) break;
			}
		}
		if (imax >= minsc) {
			if (n_b == 0 || (i32)(u32)tm.brow[n_b - 1] + 1 != i) { if (n_b < tm.brow_cap) { tm.brow[n_b] = (u64)imax << 32 | (u32)i; ++n_b; } }
			else if ((int)(tm.brow[n_b - 1] >> 32) < imax) tm.brow[n_b - 1] = (u64)imax << 32 | (u32)i;
		}
		if (imax > gmax) {
			gmax = imax; te = i;
			for (int k = 0; k < nw; ++k) tm.hmax[(size_t)k * tm.hmax_stride] = tm.H[(size_t)k * ST];
			if (gmax + shift >= 255 || gmax >= endsc) break;
		}
	}
	r.score = gmax + shift < 255 ? gmax : 255;
	r.te = te;
	if (r.score != 255) {
		if (te < 0) r.qe = 0; // (no row ever raised the maximum: the snapshot is all zeros, the first column holds its maximum)
		else {
			int best = -1;
			for (int s = 0; s < 16; ++s)
				for (int j = 0; j < slen; ++j) {
					const u32 v = tm.hmax[(size_t)(j * 8 + (s >> 1)) * tm.hmax_stride];
					const int h = (s & 1) ? half_hi(v) : half_lo(v);
					if (h > best) { best = h; r.qe = s * slen + j; }
				}
		}
		if (n_b > 0) {
			const int d = (r.score + mmax - 1) / mmax, low = te - d, high = te + d;
			for (int k = 0; k < n_b; ++k) {
				const int e = (i32)(u32)tm.brow[k];
				if ((e < low || e > high) && (int)(tm.brow[k] >> 32) > r.score2) { r.score2 = (int)(tm.brow[k] >> 32); r.te2 = e; }
			}
		}
	}
	return r;
}

// can the thread-level kernel take this task?
DEV bool tsw_takes(const MemOpt& opt, int l_ms, int tlen, int capw, int brow_cap)
{
	return l_ms > 0 && l_ms * opt.a < 250 && ((l_ms + 15) >> 4) * 8 <= capw && tlen <= brow_cap;
}
