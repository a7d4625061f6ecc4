// annot.h -- junction annotation of the index (mm_idx_t::I): BED reader and per-window junction flags, host side.
// Restates mm_idx_bed_read_core / mm_idx_bed_read_merge / mm_idx_bed_junc (index.c:672-826). The two structs keep the reference's
// layout, so an mm_idx_t annotated by either implementation can be read by the other.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cctype>
#include <string>
#include <vector>
#include <algorithm>
#include <zlib.h>

typedef struct { // index.c:35-38
	int32_t st, en, cnt;
	int32_t score:30, strand:2;
} mm_idx_intv1_t;

struct mm_idx_intv_s { // index.c:40-43
	int32_t n, m;
	mm_idx_intv1_t *a;
};

// one BED line -> intervals (BED12 with read_junc: the introns between the blocks), appended to I[id] (index.c:693-757)
template<class Name2Id> static inline void mmx_bed_line(char *line, int read_junc, int min_sc, Name2Id name2id, mm_idx_intv_s *I)
{
	mm_idx_intv1_t t = {-1, -1, -1, -1, 0};
	char *p, *q, *bl = 0, *bs = 0;
	int32_t i, id = -1, n_blk = 0;
	for (p = q = line, i = 0;; ++p) {
		if (*p == 0 || *p == '\t') {
			const int32_t c = *p;
			*p = 0;
			if (i == 0) { id = name2id(q); if (id < 0) break; }
			else if (i == 1) { t.st = atol(q); if (t.st < 0) break; }
			else if (i == 2) { t.en = atol(q); if (t.en < 0) break; }
			else if (i == 4) t.score = *q >= '0' && *q <= '9'? atol(q) : -1;
			else if (i == 5) t.strand = *q == '+'? 1 : *q == '-'? -1 : 0;
			else if (i == 9) { if (!isdigit((unsigned char)*q)) break; n_blk = atol(q); }
			else if (i == 10) bl = q;
			else if (i == 11) { bs = q; break; }
			if (c == 0) break;
			++i, q = p + 1;
		}
	}
	if (id < 0 || t.st < 0 || t.st >= t.en) return;
	if (min_sc > 0 && t.score < min_sc) return;
	mm_idx_intv_s *r = &I[id];
	auto push = [&](const mm_idx_intv1_t &s) {
		if (r->n == r->m) {
			r->m = r->m? r->m + (r->m >> 1) : 16;
			r->a = (mm_idx_intv1_t*)realloc(r->a, sizeof(*r->a) * r->m);
		}
		r->a[r->n++] = s;
	};
	if (i >= 11 && read_junc) { // BED12
		int32_t st, sz, en;
		st = strtol(bs, &bs, 10); ++bs;
		sz = strtol(bl, &bl, 10); ++bl;
		en = t.st + st + sz;
		for (i = 1; i < n_blk; ++i) {
			mm_idx_intv1_t s = t;
			st = strtol(bs, &bs, 10); ++bs;
			sz = strtol(bl, &bl, 10); ++bl;
			s.st = en, s.en = t.st + st;
			en = t.st + st + sz;
			if (s.en > s.st) push(s);
		}
	} else push(t);
}

// index.c:682-794: read, sort by (st, en), merge identical intervals (cnt = multiplicity). Returns NULL if the file cannot be opened.
template<class Name2Id> static inline mm_idx_intv_s *mmx_bed_read(const char *fn, uint32_t n_seq, int read_junc, int min_sc, Name2Id name2id, long *n_raw, long *n_merged)
{
	gzFile fp = fn && strcmp(fn, "-")? gzopen(fn, "r") : gzdopen(0, "r");
	if (fp == 0) return 0;
	mm_idx_intv_s *I = (mm_idx_intv_s*)calloc(n_seq, sizeof(mm_idx_intv_s));
	std::string line;
	int c;
	bool any = false;
	auto flush = [&]() { std::vector<char> buf(line.begin(), line.end()); buf.push_back(0); mmx_bed_line(buf.data(), read_junc, min_sc, name2id, I); line.clear(); any = false; };
	while ((c = gzgetc(fp)) >= 0) {
		if (c == '\n') flush();
		else line += (char)c, any = true;
	}
	if (any) flush();
	gzclose(fp);
	long n = 0, n0 = 0;
	for (uint32_t i = 0; i < n_seq; ++i) {
		mm_idx_intv_s *v = &I[i];
		n0 += v->n;
		// (st, en) order; the fields of merged duplicates come from the first of each group, which for a stable sort is the
		// first in file order -- only score could differ and nothing on the mapping path reads it
		std::stable_sort(v->a, v->a + v->n, [](const mm_idx_intv1_t &x, const mm_idx_intv1_t &y) { return x.st != y.st? x.st < y.st : x.en < y.en; });
		int32_t j, j0, k;
		for (j = 1, j0 = 0, k = 0; j <= v->n; ++j)
			if (j == v->n || v->a[j].st != v->a[j0].st || v->a[j].en != v->a[j0].en) {
				v->a[k] = v->a[j0];
				v->a[k++].cnt = j - j0;
				j0 = j;
			}
		v->a = (mm_idx_intv1_t*)realloc(v->a, sizeof(mm_idx_intv1_t) * (k > 0? k : 1));
		v->n = v->m = k;
		n += k;
	}
	if (n_raw) *n_raw = n0;
	if (n_merged) *n_merged = n;
	return I;
}

// index.c:802-826: bit 1/2 = first/last base of a '+' intron, 8/4 = first/last base of a '-' intron; only introns that lie
// entirely inside [st,en) are marked
static inline int mmx_bed_junc(const mm_idx_intv_s *I, int32_t n_seq, int32_t ctg, int32_t st, int32_t en, uint8_t *s)
{
	memset(s, 0, en - st);
	if (I == 0 || ctg < 0 || ctg >= n_seq) return -1;
	const mm_idx_intv_s *r = &I[ctg];
	int32_t left = 0, right = r->n;
	while (right > left) {
		const int32_t mid = left + ((right - left) >> 1);
		if (r->a[mid].st >= st) right = mid;
		else left = mid + 1;
	}
	for (int32_t i = left; i < r->n; ++i) {
		if (r->a[i].st >= en) break; // sorted by st: nothing further can lie inside the window
		if (st <= r->a[i].st && en >= r->a[i].en && r->a[i].strand != 0) {
			if (r->a[i].strand > 0) s[r->a[i].st - st] |= 1, s[r->a[i].en - 1 - st] |= 2;
			else s[r->a[i].st - st] |= 8, s[r->a[i].en - 1 - st] |= 4;
		}
	}
	return left;
}
