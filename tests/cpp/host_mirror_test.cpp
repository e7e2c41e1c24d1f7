// host_mirror_test.cpp -- exercises the C++ host mirror (recon_b200.hpp) the way recon.cpp's inner
// loop does (recon.cpp:65-119) and dumps the resulting point rows so the Python test can compare
// them with the oracle.  Usage: host_mirror_test <in.bin> <out.bin>
//   in.bin : int32 W,H,V,F,S | V*4 f32 vertices | F*3 i32 faces | 16 f32 main cam | S*16 f32 side cams
//            | W*H u8 main frame | S * W*H u8 side frames
//   out.bin: int32 M | M*7 f32 rows | int32 K | K*4 f32 filtered points | K*3 f32 filtered normals
//            (hint.filterPoints(points, normals) of recon.cpp:125 with radius = (bbox diagonal / 100)^2)
// Without a GPU it must fail loudly (exit code 3, message on stderr) -- there is no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mesh_reconstruction_b200/csrc/recon_b200.hpp"

using namespace mr;

static void rd(FILE *f, void *p, size_t n)
{
    if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("in"); return 2; }
    int32_t hdr[5];
    rd(f, hdr, sizeof(hdr));
    int W = hdr[0], H = hdr[1], V = hdr[2], F = hdr[3], S = hdr[4];
    Mat vertices(V, 4, F32), faces(F, 3, S32), mainCam(4, 4, F32), original(H, W, U8);
    rd(f, vertices.data(), (size_t)V * 16);
    rd(f, faces.data(), (size_t)F * 12);
    rd(f, mainCam.data(), 64);
    std::vector<Mat> sideCams, sideFrames;
    for (int i = 0; i < S; i++) { Mat c(4, 4, F32); rd(f, c.data(), 64); sideCams.push_back(c); }
    rd(f, original.data(), (size_t)W * H);
    for (int i = 0; i < S; i++) { Mat fr(H, W, U8); rd(f, fr.data(), (size_t)W * H); sideFrames.push_back(fr); }
    fclose(f);
    try {
        Render *render = spawnRender(W, H);                                  // recon.cpp:21
        render->loadMesh(Mesh(vertices, faces));                            // recon.cpp:42
        Mat depth = render->depth(mainCam);                                 // recon.cpp:70
        MatList flows, cameras;
        for (int i = 0; i < S; i++) {                                        // recon.cpp:81
            Mat projectedImage = render->projected(mainCam, sideFrames[i], sideCams[i]);   // recon.cpp:85
            projectedImage = mixBackground(projectedImage, original, depth);               // recon.cpp:86
            Mat flow = calculateFlow(original, projectedImage, false);                     // recon.cpp:89
            flows.push_back(flow);
            cameras.push_back(sideCams[i]);
        }
        Mat tri = triangulatePixels(flows, mainCam, cameras, depth);         // recon.cpp:114
        delete render;
        // recon.cpp:115-116,125: points / normals of the cloud, then hint.filterPoints(points, normals)
        int32_t M = tri.rows;
        Mat points(M, 4, F32), normals(M, 3, F32);
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int i = 0; i < M; i++) {
            const float *r = tri.ptr<float>(i);
            std::memcpy(points.ptr<float>(i), r, 16);
            std::memcpy(normals.ptr<float>(i), r + 4, 12);
            for (int k = 0; k < 3; k++) {
                float v = r[k] / r[3];
                if (v == v) { lo[k] = v < lo[k] ? v : lo[k]; hi[k] = v > hi[k] ? v : hi[k]; }
            }
        }
        float diag2 = 0;
        for (int k = 0; k < 3; k++) diag2 += (hi[k] - lo[k]) * (hi[k] - lo[k]);
        const float radius = diag2 / 10000.f;
        filterPoints(points, normals, radius);
        FILE *o = fopen(argv[2], "wb");
        fwrite(&M, 4, 1, o);
        fwrite(tri.data(), 1, (size_t)M * 28, o);
        int32_t K = points.rows;
        fwrite(&K, 4, 1, o);
        fwrite(&radius, 4, 1, o);
        fwrite(points.data(), 1, (size_t)K * 16, o);
        fwrite(normals.data(), 1, (size_t)K * 12, o);
        fclose(o);
        printf("host mirror: %d points, %d after filterPoints\n", M, K);
    } catch (const mr::Error &e) {
        fprintf(stderr, "mr::Error %d: %s\n", e.code, e.what());
        return 3;
    }
    return 0;
}
