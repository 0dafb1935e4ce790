// util.cuh -- reference-coordinate helpers, the tie-order-exact introsort, the mixing hash.
// Restates bns_depos / bns_pos2rid / bns_intv2rid / _get_pac, klib ks_introsort and hash_64 of
// un-vendored lh3/bwa (SURVEY.md App. B "Conventions", B.15).
#pragma once
#include "types.cuh"

DEV i64 ref_depos(const DevIndex& x, i64 pos, int& is_rev) { is_rev = pos >= x.l_pac; return is_rev ? (x.l_pac << 1) - 1 - pos : pos; }

DEV int ref_pos2rid(const DevIndex& x, i64 pos_f)
{
	if (pos_f >= x.l_pac) return -1;
	int left = 0, mid = 0, right = x.n_seqs;
	while (left < right) {
		mid = (left + right) >> 1;
		if (pos_f >= x.ctg_off[mid]) {
			if (mid == x.n_seqs - 1) break;
			if (pos_f < x.ctg_off[mid + 1]) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}

// (ctg_off has n_seqs + 1 entries, the last one l_pac: "pos_f lies in contig r" is one pair of comparisons)
DEV bool ref_pos_in_rid(const DevIndex& x, i64 pos_f, int r) { return pos_f >= x.ctg_off[r] && pos_f < x.ctg_off[r + 1]; }

// hint: a contig to try before searching (the previous seed's: the seeds of a read mostly share one); -1: none
DEV int ref_intv2rid(const DevIndex& x, i64 rb, i64 re, int hint = -1)
{
	int is_rev;
	if (rb < x.l_pac && re > x.l_pac) return -2;
	const i64 pb = ref_depos(x, rb, is_rev);
	const int rid_b = (hint >= 0 && ref_pos_in_rid(x, pb, hint)) ? hint : ref_pos2rid(x, pb);
	if (!(rb < re)) return rid_b;
	return ref_pos_in_rid(x, ref_depos(x, re - 1, is_rev), rid_b) ? rid_b : -1; // (the end in another contig: the search would return that one)
}

DEV int pac_base(const u8* pac, i64 l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

// base of the doubled (forward + reverse-complement) text at position p in [0, 2*l_pac)
DEV int ref_base2(const DevIndex& x, i64 p) { return p < x.l_pac ? pac_base(x.pac, p) : 3 - pac_base(x.pac, (x.l_pac << 1) - 1 - p); }

DEV u64 hash64(u64 key)
{
	key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
	key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
	return key;
}

// ---- klib introsort, control flow restated exactly (hazard H4: the permutation of EQUAL keys is observable) ----
// Sorts an array of T with a strict-weak `lt`.  Median-of-three with klib's comparison order, a Hoare scan that
// never examines *s, sub-ranges of <= 16 elements left for one final insertion sort, comb sort on depth overflow.
template <class T, class LT>
DEV void ks_insertsort(T* s, T* t, LT lt)
{
	for (T* i = s + 1; i < t; ++i)
		for (T* j = i; j > s && lt(*j, *(j - 1)); --j) { T tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

template <class T, class LT>
DEV void ks_combsort(size_t n, T* a, LT lt)
{
	const double shrink = 1.2473309501039786540366528676643;
	int swapped;
	size_t gap = n;
	do {
		if (gap > 2) { gap = (size_t)((double)gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
		swapped = 0;
		for (T* i = a; i < a + n - gap; ++i) {
			T* j = i + gap;
			if (lt(*j, *i)) { T tmp = *i; *i = *j; *j = tmp; swapped = 1; }
		}
	} while (swapped || gap > 2);
	if (gap != 1) ks_insertsort(a, a + n, lt);
}

template <class T, class LT>
DEV void ks_introsort(size_t n, T* a, LT lt)
{
	struct Frame { T* l; T* r; int depth; };
	Frame stack[64];
	Frame* top = stack;
	int d;
	T rp, tmp;
	T *s, *t, *i, *j, *k;
	if (n < 1) return;
	if (n == 2) { if (lt(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; }
	for (d = 2; ((size_t)1 << d) < n; ++d) {}
	s = a; t = a + (n - 1); d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) { ks_combsort((size_t)(t - s) + 1, s, lt); t = s; continue; }
			i = s; j = t; k = i + ((j - i) >> 1) + 1;
			if (lt(*k, *i)) { if (lt(*k, *j)) k = j; }
			else k = lt(*j, *i) ? i : j;
			rp = *k;
			if (k != t) { tmp = *k; *k = *t; *t = tmp; }
			for (;;) {
				do ++i; while (lt(*i, rp));
				do --j; while (i <= j && lt(rp, *j));
				if (j <= i) break;
				tmp = *i; *i = *j; *j = tmp;
			}
			tmp = *i; *i = *t; *t = tmp;
			if (i - s > t - i) {
				if (i - s > 16) { top->l = s; top->r = i - 1; top->depth = d; ++top; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { top->l = i + 1; top->r = t; top->depth = d; ++top; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == stack) { ks_insertsort(a, a + n, lt); return; }
			--top; s = top->l; t = top->r; d = top->depth;
		}
	}
}
