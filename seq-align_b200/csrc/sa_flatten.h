/*
 * sa_flatten.h -- host side: turn a reference scoring_t (256x256 swap table,
 * wildcard set, case folding, match/mismatch fallback) into the dense
 * ncodes x ncodes table the kernels index, for the characters that actually
 * occur in a batch.
 *
 * Replaces the per-cell call of scoring_lookup() (reference
 * src/alignment.c:98 -> src/alignment_scoring.c:133-182) by one lookup per
 * pair of distinct characters per batch.  The lookup rules are restated
 * here, not linked from the host library, so the engine has no dependency
 * on libalign's symbols.
 */
#ifndef SA_FLATTEN_H
#define SA_FLATTEN_H

#include <ctype.h>
#include <limits.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#include "alignment_scoring.h"

namespace sa {

struct FlatTable {
  int ncodes = 0;
  uint8_t lut[256];              /* raw byte -> code                       */
  uint8_t rep[256];              /* code -> (folded) character             */
  std::vector<int32_t> sub;      /* [cb*ncodes + ca]                       */
  std::vector<uint8_t> forbid;   /* no_mismatches && !is_match             */
  std::vector<uint8_t> unknown;  /* lookup would exit(1)                   */
  bool any_unknown = false;
  int min_sub = 0, max_sub = 0;  /* over the known entries                 */
};

struct LookupResult { int score; bool is_match; bool unknown; };

/* wildcard rule: the smaller wildcard score of the two characters wins;
 * INT_MAX doubles as "none" (alignment_scoring.c:115-129) */
static inline bool flat_wildcard(const scoring_t *s, unsigned a, unsigned b, int *score)
{
  int best = INT_MAX;
  if(get_wildcard_bit(s, a)) best = s->wildscores[a];
  if(get_wildcard_bit(s, b) && s->wildscores[b] < best) best = s->wildscores[b];
  if(best != INT_MAX) { *score = best; return true; }
  *score = 0;
  return false;
}

/* a, b already case-folded if the model is case-insensitive */
static inline LookupResult flat_lookup(const scoring_t *s, unsigned a, unsigned b)
{
  LookupResult r;
  r.unknown = false;
  r.is_match = (a == b);
  if(s->no_mismatches && !r.is_match) {           /* :148-153 */
    r.is_match = flat_wildcard(s, a, b, &r.score);
    return r;
  }
  if(get_swap_bit(s, a, b)) {                     /* :156-160 */
    r.score = s->swap_scores[a][b];
    return r;
  }
  if(flat_wildcard(s, a, b, &r.score)) {          /* :165-169 */
    r.is_match = true;
    return r;
  }
  if(s->use_match_mismatch) {                     /* :172-176 */
    r.score = r.is_match ? s->match : s->mismatch;
    return r;
  }
  r.unknown = true;                               /* :179-181 */
  r.score = 0;
  return r;
}

static inline unsigned flat_fold(const scoring_t *s, unsigned c)
{
  return s->case_sensitive ? c : (unsigned)tolower((int)c);
}

/* present_a / present_b: 256-bit sets (4 x u64) of the bytes occurring in
 * the a-side / b-side sequences of the batch */
static inline void flatten_scoring(const scoring_t *s, const uint64_t present_a[4],
                                   const uint64_t present_b[4], FlatTable *ft)
{
  int code_of[256];
  for(int i = 0; i < 256; i++) code_of[i] = -1;
  memset(ft->lut, 0, sizeof(ft->lut));
  memset(ft->rep, 0, sizeof(ft->rep));
  ft->ncodes = 0;
  bool in_a[256] = {false}, in_b[256] = {false};
  for(unsigned c = 0; c < 256; c++) {
    const bool pa = (present_a[c >> 6] >> (c & 63)) & 1, pb = (present_b[c >> 6] >> (c & 63)) & 1;
    if(!pa && !pb) continue;
    const unsigned f = flat_fold(s, c) & 0xff;
    if(code_of[f] < 0) { code_of[f] = ft->ncodes; ft->rep[ft->ncodes] = (uint8_t)f; ft->ncodes++; }
    ft->lut[c] = (uint8_t)code_of[f];
    if(pa) in_a[code_of[f]] = true;
    if(pb) in_b[code_of[f]] = true;
  }
  if(ft->ncodes == 0) { ft->ncodes = 1; ft->rep[0] = 0; }
  const int n = ft->ncodes;
  ft->sub.assign((size_t)n * n, 0);
  ft->forbid.assign((size_t)n * n, 0);
  ft->unknown.assign((size_t)n * n, 0);
  ft->any_unknown = false;
  bool first = true;
  for(int cb = 0; cb < n; cb++)
    for(int ca = 0; ca < n; ca++) {
      const LookupResult r = flat_lookup(s, ft->rep[ca], ft->rep[cb]);
      const size_t k = (size_t)cb * n + ca;
      ft->sub[k] = r.score;
      ft->forbid[k] = (s->no_mismatches && !r.is_match) ? 1 : 0;
      if(r.unknown) {
        ft->unknown[k] = 1;
        if(in_a[ca] && in_b[cb]) ft->any_unknown = true;
        continue;
      }
      if(first) { ft->min_sub = ft->max_sub = r.score; first = false; }
      if(r.score < ft->min_sub) ft->min_sub = r.score;
      if(r.score > ft->max_sub) ft->max_sub = r.score;
    }
}

} // namespace sa

#endif
