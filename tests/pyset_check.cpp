// Host build of csrc/pyset.cuh for tests/test_pyset.py (g++, no CUDA): one C entry point.
#include "../yolo_tracking_b200/csrc/pyset.cuh"

extern "C" int pyset_difference_order_c(const short* a, int na, const short* b, int nb, int nkeys, short* out) {
    static short bufs[6 * 1024];
    static unsigned char inb[4096];
    return b200::pyset_difference_order(a, na, b, nb, nkeys, bufs, 1024, inb, out);
}
