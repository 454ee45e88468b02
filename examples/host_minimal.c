/* Smallest plain-C host of the C-ABI (include/ra_b200.h): fills an ra_config with the xuzhen_12v_geo_fix_mat values, creates a
 * handle and reports what the library says.  No torch, no C++: this is what a non-Python host links against.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -Iinclude examples/host_minimal.c -Lrelightableavatar_b200 -lra_b200 \
 *       -Wl,-rpath,$PWD/relightableavatar_b200 -o /tmp/host_minimal
 *
 * On a machine without a B200 the library refuses loudly (exit code 2 and the reason on stderr): there is no CPU fallback. */
#include <stdio.h>
#include <string.h>
#include "ra_b200.h"

int main(void) {
    ra_config c;
    ra_handle* h = NULL;
    memset(&c, 0, sizeof c);
    c.relight = 1; c.precision = RA_PRECISION_TC; c.max_rays = 1 << 17; c.n_verts = 6890; c.n_bones = 52;
    c.dist_th = 0.125f; c.blend_radius = 0.075f; c.resd_limit = 0.05f;
    c.st_iter = 16; c.st_tan_i = 1000.f; c.st_relax = 0.f; c.st_offset = 0.02f; c.st_eps = 1e-8f; c.st_skip = 1;
    c.lv_iter = 4; c.lv_offset = 0.01f; c.lv_relax = 0.f; c.lv_near = 0.02f; c.lv_dist_th = 0.125f;
    c.env_r = 10.f; c.bbox_margin = 0.25f; c.render_chunk = 65536; c.n_samples = 3; c.surf_sample_range = 0.005f;
    c.fresnel_f0 = 0.02f; c.albedo_slope = 1.f; c.albedo_bias = 0.f; c.rough_slope = 0.9f; c.rough_bias = 0.09f;
    c.albedo_multiplier = 1.f; c.shading_albedo = 0.8f; c.env_h = 16; c.env_w = 32; c.vol_samples = 128;
    c.clip_near = 0.02f; c.clip_far = 10.f; c.tonemapping = 1;
    if (ra_create(&h, &c) != 0) {
        fprintf(stderr, "ra_create failed: %s\n", ra_last_error(h));
        ra_destroy(h);
        return 2;
    }
    printf("ra_b200 handle created (sizeof(ra_config) = %d bytes)\n", (int)sizeof c);
    ra_destroy(h);
    return 0;
}
