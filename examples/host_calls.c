/* host_calls.c -- the C ABI (include/qgmap.h) used from plain C99, the language of a MEX gateway.
 * Exercises the host-side entry points (no GPU needed) and shows the error contract of a compute call on a box without a CUDA
 * device.  Built and run by tests/test_abi.py::test_header_is_plain_c_and_usable:
 *     gcc -std=c99 -pedantic -Wall -Wextra -Werror -Iinclude examples/host_calls.c -Lgqmap-opticalflow_b200 -lqgmap -lm
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "qgmap.h"

int main(void)
{
    double x[5], w[5], y[3] = {0.5, 0.8, -0.1}, p[3], sum = 0.0;
    int i;
    if (qgmap_version() < 100) return 1;
    /* GaussHermite_2(5): weights sum to sqrt(pi), nodes antisymmetric */
    if (qgmap_gauss_hermite(5, x, w) != QGMAP_OK) return 2;
    for (i = 0; i < 5; ++i) sum += w[i];
    if (fabs(sum - sqrt(3.14159265358979323846)) > 1e-14 || fabs(x[0] + x[4]) > 1e-14 || fabs(x[2]) > 1e-14) return 3;
    /* projsplx([0.5 0.8 -0.1]) = [0.35 0.65 0] */
    if (qgmap_projsplx(y, 3, p) != QGMAP_OK || fabs(p[0] - 0.35) > 1e-15 || fabs(p[1] - 0.65) > 1e-15 || p[2] != 0.0) return 4;
    {   /* flowToColor on a 2 x 2 field (column-major M x N x 2): unknown-flow marker zeroed, range of the rest */
        double flow[8] = {1.0, 0.0, -2.0, 1e10, 0.5, 0.0, 0.25, 0.0}, flo[8], range[4];
        unsigned char img[12], unk[4];
        if (qgmap_flow_to_color(flow, 2, 2, -1.0, img, flo, range, unk) != QGMAP_OK) return 5;
        if (unk[3] != 1 || unk[0] != 0 || flo[3] != 0.0 || range[0] != -2.0 || range[1] != 1.0 || range[2] != 0.0 || range[3] != 0.5) return 6;
        if (img[1] != 255 || img[5] != 255 || img[9] != 255) return 7;                    /* zero flow is white */
    }
    {   /* configuration defaults are the constants of the .m files; a compute call reports why it cannot run */
        qgmap_config cfg;
        qgmap_handle *h = NULL;
        double I[64];
        int rc;
        memset(I, 0, sizeof I);
        if (qgmap_config_defaults(&cfg, QGMAP_VARIANT_FULL) != QGMAP_OK || cfg.sigma_max != 23.0 || cfg.step_tau != 8000.0) return 8;
        rc = qgmap_create(&cfg, I, I, 8, 8, &h);
        if (rc == QGMAP_OK) { qgmap_destroy(h); printf("create ok (CUDA device present)\n"); }
        else if (rc == QGMAP_ERR_CUDA && strstr(qgmap_last_error(NULL), "no CPU fallback")) printf("create refused: %s\n", qgmap_status_string(rc));
        else return 9;
    }
    printf("host calls ok\n");
    return 0;
}
