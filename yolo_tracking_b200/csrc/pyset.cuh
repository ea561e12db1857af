// Iteration order of `list(set(a) - set(b))` for small non-negative ints, as CPython 3.12 produces it (Objects/setobject.c).
//
// StrongSORT's matching cascade returns its unmatched tracks as `list(set(track_indices) - set(k for k, _ in matches))`
// (boxmot/trackers/strongsort/sort/linear_assignment.py:141).  The order of that list is the slot order of a CPython
// hash table; it decides the row order of the IoU round's cost matrix, hence - through scipy's behaviour on the tied,
// clipped costs - which detections stay unmatched in which order, hence the ids of new tracks.  The batched StrongSORT
// step therefore restates the three set operations the expression performs:
//   set(iterable)      keys added one by one (set_add_entry: linear probes of 9, then the perturbed jump; the table is
//                      rebuilt at four times the fill as soon as fill * 5 >= mask * 3, entries re-inserted in slot order)
//   a - b              len(a) >> 2 > len(b): copy of a (set_merge: same-size tables are copied slot for slot, otherwise a
//                      clean re-insertion into a table sized for 2 * len(a)), then every key of b is discarded (dummies);
//                      otherwise a new set that receives the keys of a, in a's slot order, that b does not hold
//   list(s)            slot order
// hash(k) == k for these keys.  Plain C++ (host and device): tests/test_pyset.py compiles this header with g++ and checks it
// against the running interpreter on random inputs.
#pragma once

#ifdef __CUDACC__
#define PYSET_HD __host__ __device__ __forceinline__
#else
#define PYSET_HD inline
#endif

namespace b200 {

constexpr int PYSET_LINEAR_PROBES = 9;
constexpr int PYSET_MINSIZE = 8;
constexpr short PYSET_EMPTY = -1, PYSET_DUMMY = -2;

// A table of `mask + 1` shorts inside one of two caller-provided buffers (a rebuild goes to the other buffer).
struct PySet {
    short* tab;
    short* spare;
    int mask, fill, used;
};

PYSET_HD void pyset_init(PySet& s, short* buf_a, short* buf_b) {
    s.tab = buf_a; s.spare = buf_b; s.mask = PYSET_MINSIZE - 1; s.fill = 0; s.used = 0;
    for (int i = 0; i < PYSET_MINSIZE; ++i) s.tab[i] = PYSET_EMPTY;
}

// set_insert_clean: the key is known to be absent and the table has no dummies
PYSET_HD void pyset_insert_clean(short* tab, int mask, int key) {
    unsigned long long perturb = (unsigned long long)key;
    unsigned long long i = (unsigned long long)key & (unsigned long long)mask;
    while (true) {
        if (tab[i] == PYSET_EMPTY) { tab[i] = (short)key; return; }
        if (i + PYSET_LINEAR_PROBES <= (unsigned long long)mask) {
            for (int j = 1; j <= PYSET_LINEAR_PROBES; ++j)
                if (tab[i + j] == PYSET_EMPTY) { tab[i + j] = (short)key; return; }
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & (unsigned long long)mask;
    }
}

// set_table_resize: smallest power of two above `minused`, live entries re-inserted in slot order
PYSET_HD void pyset_resize(PySet& s, int minused) {
    int newsize = PYSET_MINSIZE;
    while (newsize <= minused) newsize <<= 1;
    short* nt = s.spare;
    for (int i = 0; i < newsize; ++i) nt[i] = PYSET_EMPTY;
    for (int i = 0; i <= s.mask; ++i)
        if (s.tab[i] >= 0) pyset_insert_clean(nt, newsize - 1, s.tab[i]);
    s.spare = s.tab; s.tab = nt; s.mask = newsize - 1; s.fill = s.used;
}

// set_add_entry for a key that is not in the set (the callers add distinct keys)
PYSET_HD void pyset_add(PySet& s, int key) {
    unsigned long long perturb = (unsigned long long)key;
    unsigned long long i = (unsigned long long)key & (unsigned long long)s.mask;
    long long freeslot = -1, found = -1;
    while (found < 0) {
        const int probes = (i + PYSET_LINEAR_PROBES <= (unsigned long long)s.mask) ? PYSET_LINEAR_PROBES : 0;
        for (int j = 0; j <= probes; ++j) {
            const short e = s.tab[i + j];
            if (e == PYSET_EMPTY) { found = (long long)(i + j); break; }
            if (e == PYSET_DUMMY) freeslot = (long long)(i + j);
        }
        if (found >= 0) break;
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & (unsigned long long)s.mask;
    }
    if (freeslot >= 0) { s.tab[freeslot] = (short)key; s.used++; return; }     // found_unused_or_dummy: a dummy is reused
    s.tab[found] = (short)key;
    s.fill++; s.used++;
    if (s.fill * 5 < s.mask * 3) return;
    pyset_resize(s, s.used > 50000 ? s.used * 2 : s.used * 4);
}

// set_discard_entry (the key may be absent)
PYSET_HD void pyset_discard(PySet& s, int key) {
    unsigned long long perturb = (unsigned long long)key;
    unsigned long long i = (unsigned long long)key & (unsigned long long)s.mask;
    while (true) {
        const int probes = (i + PYSET_LINEAR_PROBES <= (unsigned long long)s.mask) ? PYSET_LINEAR_PROBES : 0;
        for (int j = 0; j <= probes; ++j) {
            const short e = s.tab[i + j];
            if (e == PYSET_EMPTY) return;
            if (e == (short)key) { s.tab[i + j] = PYSET_DUMMY; s.used--; return; }
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & (unsigned long long)s.mask;
    }
}

// list(set(a) - set(b)): a[na] and b[nb] hold distinct keys in [0, nkeys) in insertion order; out receives the result,
// the return value is its length.  bufs: six tables of `cap` shorts (cap >= 4 * the largest power of two <= 2 * na,
// 1024 covers na <= 256); inb: nkeys bytes of scratch.
PYSET_HD int pyset_difference_order(const short* a, int na, const short* b, int nb, int nkeys, short* bufs, int cap,
                                    unsigned char* inb, short* out) {
    PySet A, B, R;
    pyset_init(A, bufs, bufs + cap);
    pyset_init(B, bufs + 2 * cap, bufs + 3 * cap);
    for (int i = 0; i < na; ++i) pyset_add(A, a[i]);
    for (int i = 0; i < nb; ++i) pyset_add(B, b[i]);
    for (int i = 0; i < nkeys; ++i) inb[i] = 0;
    for (int i = 0; i < nb; ++i) inb[b[i]] = 1;
    pyset_init(R, bufs + 4 * cap, bufs + 5 * cap);
    if ((A.used >> 2) > B.used) {
        // set_copy_and_difference: set_merge into the empty result, then discard b's keys in b's slot order
        if ((R.fill + A.used) * 5 >= R.mask * 3) pyset_resize(R, (R.used + A.used) * 2);
        if (R.mask == A.mask) {
            for (int i = 0; i <= A.mask; ++i) R.tab[i] = A.tab[i];
        } else {
            for (int i = 0; i <= A.mask; ++i)
                if (A.tab[i] >= 0) pyset_insert_clean(R.tab, R.mask, A.tab[i]);
        }
        R.fill = R.used = A.used;
        for (int i = 0; i <= B.mask; ++i)
            if (B.tab[i] >= 0) pyset_discard(R, B.tab[i]);
        if (R.fill - R.used > R.mask / 4) pyset_resize(R, R.used > 50000 ? R.used * 2 : R.used * 4);
    } else {
        for (int i = 0; i <= A.mask; ++i)
            if (A.tab[i] >= 0 && !inb[A.tab[i]]) pyset_add(R, A.tab[i]);
    }
    int n = 0;
    for (int i = 0; i <= R.mask; ++i)
        if (R.tab[i] >= 0) out[n++] = R.tab[i];
    return n;
}

}  // namespace b200
