// Build-time shim: pdqsort (github.com/orlp/pdqsort @ b1ef26a5, CPM-fetched by VAMP, absent here).
// pdqsort_branchless is an unstable comparison sort; std::sort is result-equivalent up to tie order.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#pragma once
#include <algorithm>
template <class It, class Cmp> inline void pdqsort_branchless(It b, It e, Cmp c) { std::sort(b, e, c); }
template <class It> inline void pdqsort_branchless(It b, It e) { std::sort(b, e); }
template <class It, class Cmp> inline void pdqsort(It b, It e, Cmp c) { std::sort(b, e, c); }
template <class It> inline void pdqsort(It b, It e) { std::sort(b, e); }
