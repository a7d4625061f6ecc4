// tests/hostshim/hostshim.cc -- TEST INFRASTRUCTURE: C wrappers around the product's host-side hit logic (minimap2_b200/csrc/hits.cc,
// a restatement of the reference's hit.c / esterr.c) so that tests/test_hostlogic_vs_ref.py can call it next to the unmodified
// reference functions (oracle/_ref/libminimap2_ref.so) on the same inputs, without a GPU. Built by the test with g++.
#include "hostlogic.h"
#include <cstdlib>
#include <cstring>

extern "C" {
int mm_verbose = 1;
int mm_dbg_flag = 0;

mm_reg1_t *hs_gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const m128 *a, int is_qstrand) { return hl_gen_regs(hash, qlen, n_u, u, a, is_qstrand); }
void hs_set_parent(float mask_level, int mask_len, int n, mm_reg1_t *r, int sub_diff, int hard_mask_level, float alt_diff_frac) { hl_set_parent(mask_level, mask_len, n, r, sub_diff, hard_mask_level, alt_diff_frac); }
void hs_select_sub(float pri_ratio, int min_diff, int best_n, int check_strand, int min_strand_sc, int *n_, mm_reg1_t *r) { hl_select_sub(pri_ratio, min_diff, best_n, check_strand, min_strand_sc, n_, r); }
void hs_hit_sort(int *n_regs, mm_reg1_t *r, float alt_diff_frac) { hl_hit_sort(n_regs, r, alt_diff_frac); }
int hs_set_sam_pri(int n, mm_reg1_t *r) { return hl_set_sam_pri(n, r); }
void hs_sync_regs(int n, mm_reg1_t *r) { hl_sync_regs(n, r); }
int hs_filter_strand_retained(int n, mm_reg1_t *r) { return hl_filter_strand_retained(n, r); }
void hs_set_mapq(int n, mm_reg1_t *r, int min_chain_sc, int match_sc, int rep_len, int is_sr, int is_splice) { hl_set_mapq(n, r, min_chain_sc, match_sc, rep_len, is_sr, is_splice); }
void hs_est_err(const mm_idx_t *mi, int qlen, int n_regs, mm_reg1_t *regs, const m128 *a, int32_t n, const uint64_t *mini_pos) { hl_est_err(mi, qlen, n_regs, regs, a, n, mini_pos); }
void hs_free(void *p) { free(p); }
}

// symmetric DUST (csrc/hits.cc hl_sdust) next to the reference's sdust(): returns the number of intervals, copies up to cap of them
extern "C" int hs_sdust(const uint8_t *seq, int l_seq, int T, int W, uint64_t *out, int cap)
{
	std::vector<uint64_t> res;
	hl_sdust(seq, l_seq, T, W, res);
	for (int i = 0; i < (int)res.size() && i < cap; ++i) out[i] = res[i];
	return (int)res.size();
}
