// Build-time shim: pdqsort (github.com/orlp/pdqsort @ b1ef26a5, CPM-fetched by VAMP, absent here).
// pdqsort_branchless is an unstable comparison sort; std::sort is result-equivalent up to tie order. refv_sort_mode() = 1 switches to
// std::stable_sort (refv_set_stable_sort): the same reference code with ties kept in their previous order -- the order the engine and the
// restatement define --, which shows that tie order is the ONLY thing an output can differ by.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#pragma once
#include <algorithm>
inline int &refv_sort_mode() { static int mode = 0; return mode; }
template <class It, class Cmp> inline void pdqsort_branchless(It b, It e, Cmp c) { if (refv_sort_mode()) std::stable_sort(b, e, c); else std::sort(b, e, c); }
template <class It> inline void pdqsort_branchless(It b, It e) { if (refv_sort_mode()) std::stable_sort(b, e); else std::sort(b, e); }
template <class It, class Cmp> inline void pdqsort(It b, It e, Cmp c) { if (refv_sort_mode()) std::stable_sort(b, e, c); else std::sort(b, e, c); }
template <class It> inline void pdqsort(It b, It e) { if (refv_sort_mode()) std::stable_sort(b, e); else std::sort(b, e); }
