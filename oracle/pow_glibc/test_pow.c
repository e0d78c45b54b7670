/* checks ss_pow_pos against this machine's libm pow() on random operands of the welfare domain and beyond */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../../sugarscape_utilicalc_b200/csrc/ss_pow.h"
static unsigned long long s = 88172645463325252ull;
static unsigned long long rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double u01(void) { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 20000000, bad = 0, skipped = 0;
    ss_pow_tables T = {SS_POW_LOG_HDR, SS_POW_LOG_TAB, SS_POW_EXP_HDR, SS_POW_EXP_TAB};
    for (long t = 0; t < n; t++) {
        double x, y;
        int mode = (int)(rnd() % 4);
        if (mode == 0) { x = (double)(1 + rnd() % 400); int m1 = 1 + rnd() % 8, m2 = 1 + rnd() % 8; y = (double)m1 / (double)(m1 + m2); }    /* integer wealth */
        else if (mode == 1) { x = 0.01 + u01() * 500.0; int m1 = 1 + rnd() % 20, m2 = 1 + rnd() % 20; y = (double)m1 / (double)(m1 + m2); }
        else if (mode == 2) { x = exp((u01() - 0.5) * 60.0); y = (u01() - 0.5) * 8.0; }
        else { x = 0.5 * (double)(1 + rnd() % 2000); y = u01(); }
        if (x == 1.0 || y == 0.0) { skipped++; continue; }
        int b = 0;
        double got = ss_pow_pos(&T, x, y, &b), want = pow(x, y);
        if (b) { skipped++; continue; }
        if (memcmp(&got, &want, 8)) { if (bad < 10) printf("MISMATCH x=%a y=%a got=%a want=%a\n", x, y, got, want); bad++; }
    }
    printf("%ld operands, %ld outside the main path, %ld mismatches\n", n, skipped, bad);
    return bad != 0;
}
