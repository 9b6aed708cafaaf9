// The slice of parasail's C interface the reference's alignment path and its unit test touch (vendor/parasail/parasail.h:
// matrix create / free, result free, traceback get / free), backed by the device alignment of the C ABI.  Scoring is the
// path's: match +2, mismatch -2 over ACGT, gap extension 1 (cluster.cpp:473-479); other values are refused.
#ifndef ISONCLUST2_COMPAT_PARASAIL_H_INCLUDED
#define ISONCLUST2_COMPAT_PARASAIL_H_INCLUDED

#include <cstring>
#include <string>

#include "icb_compat.h"

struct parasail_matrix_t {
    int match, mismatch;
    std::string alphabet;
};

struct parasail_result_t {
    int score, end_query, end_ref;
    std::string comp;  // the traceback's comparison line, computed with the alignment
};

struct parasail_traceback_t {
    char* query;
    char* comp;
    char* ref;
};

inline parasail_matrix_t* parasail_matrix_create(const char* alphabet, int match, int mismatch)
{
    return new parasail_matrix_t{match, mismatch, alphabet ? alphabet : ""};
}
inline void parasail_matrix_free(parasail_matrix_t* m) { delete m; }
inline void parasail_result_free(parasail_result_t* r) { delete r; }
inline int parasail_result_get_score(const parasail_result_t* r) { return r->score; }
inline int parasail_result_get_end_query(const parasail_result_t* r) { return r->end_query; }
inline int parasail_result_get_end_ref(const parasail_result_t* r) { return r->end_ref; }
inline int parasail_result_is_saturated(const parasail_result_t*) { return 0; }

// only the comparison line is materialised (what getAlnRatio reads); query / ref lines are left empty
inline parasail_traceback_t* parasail_result_get_traceback(parasail_result_t* result, const char*, int, const char*, int,
							   const parasail_matrix_t*, char match, char pos, char neg)
{
    parasail_traceback_t* t = new parasail_traceback_t;
    t->comp = new char[result->comp.size() + 1];
    for (size_t i = 0; i < result->comp.size(); i++) t->comp[i] = result->comp[i] == '|' ? match : (pos == neg ? pos : neg);
    t->comp[result->comp.size()] = 0;
    t->query = new char[1]{0};
    t->ref = new char[1]{0};
    return t;
}
inline void parasail_traceback_free(parasail_traceback_t* t)
{
    if (!t) return;
    delete[] t->query;
    delete[] t->comp;
    delete[] t->ref;
    delete t;
}

#endif
