/* A plain-C client of librbsim_b200.so: shows that the boundary needs nothing but include/rbsim_b200.h.
 *   cabi_client validate   -- argument validation only (no GPU needed)
 *   cabi_client sphere     -- config 1 (single sphere from rest, 2000 steps) through the host-buffer driver;
 *                             prints the final qpos / qvel (needs a GPU)
 * Built and run by tests/test_cabi.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rbsim_b200.h"

static void sphere_args(rbs_body_plane_args *a, long n_env) {
    memset(a, 0, sizeof(*a));
    a->dtype = RBS_F64;
    a->geom = RBS_GEOM_SPHERE;
    a->scheme = RBS_SCHEME_A;
    a->inertia_mode = RBS_INERTIA_GENERAL;      /* strict policy, literal world inertia */
    a->arith = RBS_ARITH_STRICT;
    a->n_env = n_env;
    a->substeps = 50;
    a->mass_u = 1.675516081914557;              /* models/sphere.xml: r = 0.2, density 50 */
    a->inertia_u[0] = a->inertia_u[1] = a->inertia_u[2] = 0.026808257310632914;
    a->size_u[0] = 0.2;
    a->restitution_u = 1.0;                     /* src/config/sim_overrides.py: single_sphere_bounce */
    a->friction_u = 0.5;
    a->plane_normal[2] = 1.0;
    a->gravity[2] = -9.8;
    a->dt = 0.009;
    a->contact_threshold = 0.0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (rbs_version() != RBS_ABI_VERSION) return 3;
    if (strcmp(argv[1], "validate") == 0) {
        rbs_body_plane_args a;
        sphere_args(&a, 4);
        a.dtype = 42;
        if (rbs_step_body_plane(&a) != RBS_EINVAL || !strstr(rbs_last_error(), "dtype")) return 10;
        sphere_args(&a, 4);
        a.substeps = 0;
        if (rbs_step_body_plane(&a) != RBS_EINVAL) return 11;
        sphere_args(&a, 4);
        a.stride = 2;
        a.state = (void *)0x10;
        if (rbs_step_body_plane(&a) != RBS_EINVAL || !strstr(rbs_last_error(), "stride")) return 12;
        sphere_args(&a, 0);
        if (rbs_step_body_plane(&a) != RBS_OK) return 13;                /* empty batch is a no-op */
        if (rbs_run_body_plane_host(&a, NULL, NULL, 10) != RBS_OK) return 14;
        sphere_args(&a, 4);
        if (rbs_run_body_plane_host(&a, NULL, NULL, 10) != RBS_EINVAL && rbs_device_count() > 0) return 15;
        printf("validate ok\n");
        return 0;
    }
    if (strcmp(argv[1], "sphere") == 0) {
        rbs_body_plane_args a;
        double qpos[7] = {0, 0, 2.0, 1, 0, 0, 0}, qvel[6] = {0, 0, 0, 0, 0, 0};
        sphere_args(&a, 1);
        int rc = rbs_run_body_plane_host(&a, qpos, qvel, 2000);
        if (rc != RBS_OK) {
            fprintf(stderr, "rbs_run_body_plane_host: %d %s\n", rc, rbs_last_error());
            return 20;
        }
        printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", qpos[0], qpos[1], qpos[2], qpos[3], qpos[4], qpos[5],
               qpos[6], qvel[2]);
        rbs_release_workspace();
        return 0;
    }
    return 2;
}
