/* hmc_order.h -- the chains of an HMC ensemble in order of trajectory length (host code, plain C).
 *
 * kHmcLeapDmma (contraction.cuh) launches only the row tiles of the chains that still run a given
 * leap-frog stage; for that the rows of its launches are the chains sorted by the number of stages
 * they take, longest first.  Used by engine_hmc.inl on the lengths it reads back from the device and
 * by the host known-answer library (host_kat.c) so that the CPU test-suite checks it.
 *
 *   steps[c]      trajectory length of chain c (< 1: no trajectory in this transition)
 *   order[r]      out: the chain in row r -- lengths descending, equal lengths in chain order, the
 *                 chains without a trajectory last (in chain order)
 *   tiles[k]      out, k = 0 .. maxSteps: row tiles of `tile` rows that hold every chain taking part in
 *                 gradient k (all chains with steps >= max(k, 1))
 *   scratch       maxSteps + 2 ints
 * Returns the sum of tiles[k]: the row tiles the maxSteps + 1 launches compute. */
#ifndef SMCMC_HMC_ORDER_H_SEEN
#define SMCMC_HMC_ORDER_H_SEEN

static inline long long smcmc_hmc_order(const int* steps, int chains, int maxSteps, int tile, int* order, int* tiles,
                                        int* scratch) {
    int* atLeast = scratch;                          /* atLeast[k] = chains with length >= k, k = 1 .. maxSteps + 1 */
    for (int k = 0; k <= maxSteps + 1; ++k) atLeast[k] = 0;
    for (int c = 0; c < chains; ++c) {
        const int st = steps[c];
        if (st >= 1) atLeast[st < maxSteps ? st : maxSteps] += 1;
    }
    for (int k = maxSteps - 1; k >= 1; --k) atLeast[k] += atLeast[k + 1];
    long long busy = 0;
    for (int k = 0; k <= maxSteps; ++k) {
        const int active = maxSteps >= 1 ? atLeast[k > 1 ? k : 1] : 0;
        tiles[k] = (active + tile - 1) / tile;
        busy += tiles[k];
    }
    if (order) {
        /* first row of each length, longest first; then reuse atLeast[] as the running positions */
        int pos = 0;
        for (int st = maxSteps; st >= 1; --st) {
            const int count = atLeast[st] - atLeast[st + 1];
            atLeast[st + 1] = pos;                   /* position of length st, kept one slot up */
            pos += count;
        }
        int tail = pos;
        for (int c = 0; c < chains; ++c) {
            const int st = steps[c];
            if (st >= 1) order[atLeast[(st < maxSteps ? st : maxSteps) + 1]++] = c;
            else order[tail++] = c;
        }
    }
    return busy;
}

#endif
