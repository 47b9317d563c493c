// TEST INFRASTRUCTURE ONLY: the host check (hostcheck.cpp) with a counter on the CP_TRACE_PRUNE hook, built twice by
// tests/test_prune_option.py (-DCP_PRUNE=1 and default) to show that the experimental pruning of certainly rejected
// line-search trials (csrc/cp_point.cuh, cp_prune_setup) changes nothing but the number of evaluations carried out.
#include <stdint.h>
static long long g_pruned = 0;
#define CP_TRACE_PRUNE() (++g_pruned)
#include "hostcheck.cpp"
extern "C" long long hostcheck_pruned_count() { const long long v = g_pruned; g_pruned = 0; return v; }
