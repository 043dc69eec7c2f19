#include <math.h>
#include <stdio.h>
int main(void) {
    long bad = 0, n = 0;
    for (int b = 1; b <= 40000; b++) {
        double db = b, rb = 1.0 / db;
        for (int a = 1; a < b; a++) {
            double da = a, q0 = da * rb, rem = fma(-db, q0, da), q = fma(rem, rb, q0);
            if (q != da / db) bad++;
            n++;
        }
    }
    printf("checked %ld quotients, mismatches %ld\n", n, bad);
    return bad != 0;
}
