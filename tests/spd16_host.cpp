// Host build of bayespy_b200/csrc/spd16.cuh for tests/test_spd16.py: one column, scratch rows contiguous.
#include "../bayespy_b200/csrc/spd16.cuh"
struct HostAcc {
    double *p;
    double ld(int row) const { return p[row]; }
    void st(int row, double v) { p[row] = v; }
};
extern "C" int spd16_host(double *rows /* [SPD16_ROWS] */, double *q, double *ld) {
    HostAcc a{rows};
    return spd16_solve_inverse(a, *q, *ld) ? 0 : 1;
}
extern "C" int spd16_rows(void) { return SPD16_ROWS; }
