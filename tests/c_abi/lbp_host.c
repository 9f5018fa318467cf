/* A host program in plain C over the C ABI of include/hsm.h -- what a non-Python caller of the path
 * binds.  Reads one case from a binary file, runs the reference's `lbp` call sequence
 * (mrf.py:146-167: loop_belief(reinit) + get_map_belief) and writes the labels and the five planes.
 *
 * input : int32 rows, cols, levels, n_iter, has_prior, mode; float32 keep32; float64 keep64;
 *         float64 left[rows*cols], right[rows*cols], prior[rows*cols] (if has_prior)
 * output: int64 labels[rows*cols]; float32 plane[5][rows*cols*levels] in the library's (row, col, level) order
 */
#include <stdio.h>
#include <stdlib.h>

#include "hsm.h"

static int fail(const char *what)
{
    fprintf(stderr, "%s: %s\n", what, hsm_last_error());
    return 1;
}

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    FILE *in = fopen(argv[1], "rb");
    if (!in) return 2;
    int head[6];
    float keep32;
    double keep64;
    if (fread(head, sizeof head, 1, in) != 1 || fread(&keep32, sizeof keep32, 1, in) != 1 ||
        fread(&keep64, sizeof keep64, 1, in) != 1)
        return 2;
    const int rows = head[0], cols = head[1], levels = head[2], n_iter = head[3], has_prior = head[4], mode = head[5];
    const size_t n = (size_t)rows * cols;
    double *left = malloc(n * sizeof *left), *right = malloc(n * sizeof *right), *prior = malloc(n * sizeof *prior);
    long long *labels = malloc(n * sizeof *labels);
    float *plane = malloc(n * levels * sizeof *plane);
    if (!left || !right || !prior || !labels || !plane) return 2;
    if (fread(left, sizeof *left, n, in) != n || fread(right, sizeof *right, n, in) != n) return 2;
    if (has_prior && fread(prior, sizeof *prior, n, in) != n) return 2;
    fclose(in);

    hsm_matcher *m = NULL;
    if (hsm_create(rows, cols, levels, 1, 0, &m) != HSM_OK) return fail("hsm_create");
    if (hsm_lbp(m, 1, left, right, HSM_F64, has_prior ? prior : NULL, HSM_F64, has_prior ? mode : HSM_PRIOR_NONE, keep32,
                keep64, n_iter, (int64_t *)labels, 0, 0) != HSM_OK)
        return fail("hsm_lbp");

    FILE *out = fopen(argv[2], "wb");
    if (!out) return 2;
    fwrite(labels, sizeof *labels, n, out);
    for (int which = 0; which < 5; ++which) {
        if (hsm_get_plane(m, 0, which, plane) != HSM_OK) return fail("hsm_get_plane");
        fwrite(plane, sizeof *plane, n * levels, out);
    }
    fclose(out);
    if (hsm_destroy(m) != HSM_OK) return fail("hsm_destroy");
    free(left); free(right); free(prior); free(labels); free(plane);
    return 0;
}
