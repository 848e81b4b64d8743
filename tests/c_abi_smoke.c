/* c_abi_smoke.c -- the C-ABI used the way a C / C++ planner would: no Python anywhere.
 *
 *   c_abi_smoke <robot.urdf> <scene.yaml> <cases.txt>
 *
 * Loads the robot and the scene with the file-level loaders (jg_robot_load / jg_scene_load), creates an engine on
 * device 0 and checks the configurations of <cases.txt> (first line: n; then per line dof values and the expected
 * verdict: 0 free, 1 colliding, -1 inside the 1 mm band = do not compare) with jg_check_host (limits + self-collision +
 * scene + plane, the filter of scenario.py:109-119).  Prints one summary line; exit code 0 iff every compared verdict
 * agrees.  tests/test_c_loaders.py builds it with gcc, feeds it the reference's golden IK rows (known negatives, G4) and
 * oracle-labelled random configurations, and runs it on the GPU box.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "jogramop_b200.h"

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s robot.urdf scene.yaml cases.txt\n", argv[0]); return 2; }
    jg_robot* robot = jg_robot_load(argv[1], NULL);
    if (!robot) { fprintf(stderr, "jg_robot_load: %s\n", jg_last_error()); return 3; }
    jg_scene* scene = jg_scene_load(argv[2], NULL);
    if (!scene) { fprintf(stderr, "jg_scene_load: %s\n", jg_last_error()); return 3; }
    jg_robot_desc rd;
    jg_scene_desc sd;
    jg_robot_describe(robot, &rd);
    jg_scene_describe(scene, &sd);
    jg_engine* eng = jg_engine_create(robot, scene, 0, NULL);
    if (!eng) { fprintf(stderr, "jg_engine_create: %s\n", jg_last_error()); return 4; }
    /* lanes: arenas a device-pointer call of several passes may use side by side (not used by the host-buffer call below) */
    if (jg_engine_set_lanes(eng, 2) != JG_OK || jg_engine_lanes(eng) != 2 || jg_engine_set_lanes(eng, 9) == JG_OK ||
        jg_engine_pass_size(eng) < 32) { fprintf(stderr, "lanes / pass size: %s\n", jg_last_error()); return 4; }
    FILE* f = fopen(argv[3], "r");
    int n = 0;
    if (!f || fscanf(f, "%d", &n) != 1 || n <= 0) { fprintf(stderr, "cannot read %s\n", argv[3]); return 2; }
    float* q = (float*)malloc(sizeof(float) * (size_t)n * rd.dof);
    int* expect = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < rd.dof; ++j)
            if (fscanf(f, "%f", &q[(size_t)i * rd.dof + j]) != 1) { fprintf(stderr, "short case file\n"); return 2; }
        if (fscanf(f, "%d", &expect[i]) != 1) { fprintf(stderr, "short case file\n"); return 2; }
    }
    fclose(f);
    uint32_t* bits = (uint32_t*)calloc((size_t)(n + 31) / 32, sizeof(uint32_t));
    float* dist = (float*)malloc(sizeof(float) * (size_t)n);
    uint8_t* reason = (uint8_t*)malloc((size_t)n);
    int rc = jg_check_host(eng, q, n, rd.dof, JG_SCENE | JG_PLANE | JG_SELF | JG_LIMITS, bits, dist, reason);
    if (rc != JG_OK) { fprintf(stderr, "jg_check_host: %s\n", jg_last_error()); return 5; }
    int compared = 0, wrong = 0, colliding = 0;
    for (int i = 0; i < n; ++i) {
        const int got = (int)((bits[i >> 5] >> (i & 31)) & 1u);
        colliding += got;
        if (expect[i] < 0) continue;
        ++compared;
        if (got != expect[i]) {
            if (wrong < 5) fprintf(stderr, "case %d: got %d, expected %d (distance %g, reason %d)\n", i, got, expect[i], dist[i], (int)reason[i]);
            ++wrong;
        }
    }
    jg_stats st;
    jg_get_stats(eng, &st);
    printf("c_abi_smoke: links %d shapes %d dof %d self-pairs %d ee %d | pieces %d bodies %d | %d cases, %d colliding, "
           "%d compared, %d wrong | kernel launches %lld\n", rd.n_links, rd.n_shapes, rd.dof, rd.n_self_pairs,
           jg_robot_link_index(robot, "panda_grasptarget"), sd.n_pieces, sd.n_bodies, n, colliding, compared, wrong,
           (long long)st.kernel_launches);
    jg_engine_free(eng);
    jg_scene_free(scene);
    jg_robot_free(robot);
    free(q); free(expect); free(bits); free(dist); free(reason);
    return wrong == 0 ? 0 : 1;
}
