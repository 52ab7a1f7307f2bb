/* abi_check.c -- the C ABI from plain C99: include/dsrt.h, dsrt_types.h and dsrt_legacy.h compile without a C++ or
 * CUDA compiler, the record sizes are the reference's, and the host-side entry points work without a GPU.
 * Built and run by tests/test_host.py::test_headers_are_plain_c_and_link_from_c. */
#include <stdio.h>
#include <string.h>

#include "dsrt.h"
#include "dsrt_legacy.h"

int main(int argc, char** argv) {
    if (sizeof(DsrtTriangle) != 116 || sizeof(DsrtMaterial) != 48 || sizeof(DsrtSphere) != 24 ||
        sizeof(DsrtTextureHeader) != 12 || sizeof(DsrtCamera) != 104 || sizeof(DsrtRenderParams) != 48 ||
        sizeof(DsrtLegacyScene) != 384) {
        printf("record sizes differ from the reference's GPUScene ABI\n");
        return 1;
    }
    printf("%s\n", dsrt_version());
    if (argc > 1) {
        double poses[9 * 4];
        const int n = dsrt_read_pose_file(argv[1], poses, 4);
        float cam[3], sun[3];
        double sep = 0.0;
        DsrtCamera c;
        if (n < 4) { printf("pose file: %d poses\n", n); return 2; }
        dsrt_pose_to_model_frame(poses, cam, sun, &sep);
        dsrt_camera_look_at_origin(cam, 40.0f, 800, 450, 1000, 50, &c);
        printf("poses %d sep %.6f cam %.9g %.9g %.9g llc %.9g\n", n, sep, cam[0], cam[1], cam[2], c.lower_left_corner.x);
    }
    {
        dsrt_ctx* ctx = NULL;
        const int rc = dsrt_create(&ctx, 0);
        if (rc == DSRT_OK) { printf("device ok\n"); dsrt_destroy(ctx); }
        else printf("no device: rc %d: %s\n", rc, dsrt_last_error(NULL));
    }
    return 0;
}
