// recon_b200.hpp -- C++ host-side mirror of the reference's interface for the hot path
// (recon.hpp:18-25,40-55,93-100), layered on the C ABI of libmeshrecon_b200.so.
//
// The reference's types are cv::Mat based; OpenCV's C++ headers are not available in this
// build image, so this header carries a minimal row-major matrix type `mr::Mat` with the
// handful of members the path uses (rows, cols, channels, typed data).  Where OpenCV IS
// available (the reference's own build), define MR_WITH_OPENCV before including this file:
// the same functions are then also provided on cv::Mat, with the reference's exact names and
// signatures, so that render_cuda.cpp / flow.cpp / util.cpp of INTEGRATION.md are one-liners.
//
// Error behaviour mirrors the reference: it aborts via assert/exit on misuse
// (render_glx.cpp:66,231; util.cpp:368-369,442); here every failing C-ABI call throws
// mr::Error carrying mr_last_error(), and the asserts on channel counts are kept.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/meshrecon_b200.h"

namespace mr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

enum Depth { U8 = 0, S32 = 4, F32 = 5 };  // same numeric values as CV_8U / CV_32S / CV_32F

// Minimal dense row-major matrix (continuous storage, shared ownership like cv::Mat).
struct Mat {
    int rows = 0, cols = 0, chans = 1, depth = F32;
    std::shared_ptr<std::vector<uint8_t>> store;
    Mat() {}
    Mat(int r, int c, int d, int ch = 1) : rows(r), cols(c), chans(ch), depth(d), store(std::make_shared<std::vector<uint8_t>>((size_t)r * c * ch * elem1(d))) {}
    static size_t elem1(int d) { return d == U8 ? 1 : 4; }
    int channels() const { return chans; }
    bool empty() const { return rows == 0 || cols == 0; }
    template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(store->data()) + (size_t)r * cols * chans; }
    template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(store->data()) + (size_t)r * cols * chans; }
    uint8_t *data() { return store ? store->data() : nullptr; }
    const uint8_t *data() const { return store ? store->data() : nullptr; }
    Mat clone() const { Mat m(rows, cols, depth, chans); if (store) std::memcpy(m.data(), data(), store->size()); return m; }
};

typedef struct Mesh {  // recon.hpp:19-21
    Mat vertices, faces;
    Mesh(Mat v, Mat f) : vertices(v), faces(f) {}
} Mesh;
typedef std::list<Mat> MatList;  // recon.hpp:25
const float backgroundDepth = MR_BACKGROUND_DEPTH;  // recon.hpp:30

namespace detail {
inline void check(mr_context *ctx, int rc)
{
    if (rc != MR_OK) throw Error(rc, mr_last_error(ctx));
}
// one context per (thread, device, width, height), shared by Render and the free functions: calls on a context are
// serialised by its owner (meshrecon_b200.h), so every host thread gets its own
inline mr_context *ctx_for(int w, int h, int device = 0)
{
    struct Key { int d, w, h; mr_context *c; };
    thread_local std::vector<Key> mine;
    for (auto &k : mine) if (k.d == device && k.w == w && k.h == h) return k.c;
    mr_context *c = nullptr;
    int rc = mr_create(&c, device, w, h);
    if (rc != MR_OK) throw Error(rc, mr_last_error(nullptr));
    mine.push_back({device, w, h, c});
    return c;
}
}  // namespace detail

// == render_glx.cpp replacement: class Render (recon.hpp:93-99) ==
class Render {
  public:
    virtual ~Render() {}
    virtual void loadMesh(const Mesh) = 0;
    virtual Mat projected(const Mat camera, const Mat frame, const Mat projector) = 0;
    virtual Mat depth(const Mat camera) const = 0;
};

class RenderCUDA : public Render {
  public:
    RenderCUDA(int width, int height, int device = 0) : w(width), h(height), ctx(detail::ctx_for(width, height, device)) {}
    void loadMesh(const Mesh mesh) override
    {
        assert(mesh.vertices.cols == 4 && mesh.faces.cols == 3);  // render_glx.cpp:231
        detail::check(ctx, mr_load_mesh(ctx, mesh.vertices.ptr<float>(), mesh.vertices.rows, mesh.faces.ptr<int32_t>(), mesh.faces.rows));
    }
    Mat projected(const Mat camera, const Mat frame, const Mat projector) override
    {
        assert(frame.channels() == 1);  // render_glx.cpp:66
        Mat out(h, w, U8, 3);
        detail::check(ctx, mr_projected(ctx, camera.ptr<float>(), frame.ptr<uint8_t>(), projector.ptr<float>(), out.ptr<uint8_t>()));
        return out;
    }
    Mat depth(const Mat camera) const override
    {
        Mat out(h, w, F32, 1);
        detail::check(ctx, mr_depth(ctx, camera.ptr<float>(), out.ptr<float>()));
        return out;
    }
    mr_context *context() const { return ctx; }

  private:
    int w, h;
    mr_context *ctx;
};

// spawnRender(Heuristic hint) (recon.hpp:100): the size comes from hint.renderSize().
inline Render *spawnRender(int width, int height, int device = 0) { return new RenderCUDA(width, height, device); }

// == flow.cpp ==  Mat calculateFlow(const Mat prev, const Mat next, bool useFarneback)   recon.hpp:40
inline Mat calculateFlow(const Mat prev, const Mat next, bool useFarneback)
{
    mr_context *c = detail::ctx_for(prev.cols, prev.rows);
    Mat out(prev.rows, prev.cols, F32, 4);
    detail::check(c, mr_calculate_flow(c, prev.ptr<uint8_t>(), next.ptr<uint8_t>(), useFarneback ? 1 : 0, out.ptr<float>()));
    return out;
}

// == util.cpp ==
inline Mat mixBackground(const Mat image, const Mat background, Mat &depth)  // recon.hpp:49 (depth is in/out)
{
    assert(image.channels() == 3);       // util.cpp:368
    assert(background.channels() == 1);  // util.cpp:369
    mr_context *c = detail::ctx_for(depth.cols, depth.rows);
    Mat out(depth.rows, depth.cols, U8, 1);
    detail::check(c, mr_mix_background(c, image.ptr<uint8_t>(), background.ptr<uint8_t>(), depth.ptr<float>(), out.ptr<uint8_t>()));
    return out;
}
inline Mat flowRemap(const Mat flow, const Mat image)  // recon.hpp:50
{
    mr_context *c = detail::ctx_for(image.cols, image.rows);
    Mat out(image.rows, image.cols, U8, 1);
    detail::check(c, mr_flow_remap(c, flow.ptr<float>(), flow.channels(), image.ptr<uint8_t>(), out.ptr<uint8_t>()));
    return out;
}
inline Mat compare(const Mat prev, const Mat next)  // recon.hpp:45
{
    mr_context *c = detail::ctx_for(prev.cols, prev.rows);
    Mat out(prev.rows, prev.cols, F32, 1);
    detail::check(c, mr_compare(c, prev.ptr<uint8_t>(), next.ptr<uint8_t>(), out.ptr<float>()));
    return out;
}
inline Mat imageGradient(const Mat image)  // recon.hpp:55
{
    mr_context *c = detail::ctx_for(image.cols, image.rows);
    Mat out(image.rows, image.cols, F32, 2);
    detail::check(c, mr_image_gradient(c, image.ptr<float>(), out.ptr<float>()));
    return out;
}
inline Mat extractCameraCenter(const Mat camera)  // recon.hpp:43 (returned dehomogenised, 3x1)
{
    Mat out(3, 1, F32, 1);
    int rc = mr_extract_camera_center(camera.ptr<float>(), out.ptr<float>());
    if (rc != MR_OK) throw Error(rc, "mr_extract_camera_center");
    return out;
}
// Mat triangulatePixels(const MatList flows, const Mat mainCamera, const MatList cameras, const Mat depth)  recon.hpp:44
inline Mat triangulatePixels(const MatList flows, const Mat mainCamera, const MatList cameras, const Mat depth)
{
    assert(flows.size() == cameras.size() && !flows.empty());
    mr_context *c = detail::ctx_for(depth.cols, depth.rows);
    std::vector<const float *> fp;
    std::vector<float> cams;
    for (const Mat &f : flows) fp.push_back(f.ptr<float>());
    for (const Mat &m : cameras) cams.insert(cams.end(), m.ptr<float>(), m.ptr<float>() + 16);
    Mat all(depth.rows * depth.cols, 7, F32, 1);
    int count = 0;
    detail::check(c, mr_triangulate_pixels(c, fp.data(), (int)fp.size(), mainCamera.ptr<float>(), cams.data(), depth.ptr<float>(),
                                           all.ptr<float>(), &count));
    Mat out(count, 7, F32, 1);  // "points.resize(pixelId)" util.cpp:248
    if (count) std::memcpy(out.data(), all.data(), (size_t)count * 7 * sizeof(float));
    return out;
}


// == heuristic.cpp ==  void Heuristic::filterPoints(Mat& points, Mat& normals)   recon.hpp:115, heuristic.cpp:55-176
// `radius` is the alphaVals.back() / 4 the reference derives from its alpha shape (heuristic.cpp:63); like there it
// bounds SQUARED distances.  points (n x 4) and normals (n x 3) are replaced by the survivors, in ascending index order.
inline void filterPoints(Mat &points, Mat &normals, float radius, int device = 0)
{
    assert(points.cols == 4 && normals.cols == 3 && points.rows == normals.rows);
    mr_context *c = detail::ctx_for(16, 16, device);           // the filter does not depend on the render size
    Mat op(points.rows ? points.rows : 1, 4, F32), on(points.rows ? points.rows : 1, 3, F32);
    size_t m = 0;
    detail::check(c, mr_filter_points(c, points.ptr<float>(), normals.ptr<float>(), (size_t)points.rows, radius, op.ptr<float>(), on.ptr<float>(),
                                      nullptr, &m));
    Mat p2((int)m, 4, F32), n2((int)m, 3, F32);                // points.resize(writeIndex)  heuristic.cpp:174-175
    if (m) { std::memcpy(p2.data(), op.data(), m * 16); std::memcpy(n2.data(), on.data(), m * 12); }
    points = p2;
    normals = n2;
}

// == configuration.cpp:226-245 ==  what Configuration does to every decoded frame: cv::resize(INTER_AREA) to the render
// size (integer factor) + cv::cvtColor(CV_BGR2GRAY)
// exposure != nullptr: estimateExposure's three channel weights (B, G, R) of this frame replace the gray conversion
// (configuration.cpp:417-425; the estimate itself is host work)
inline Mat ingestFrame(const Mat bgr, int width, int height, int device = 0, const float *exposure = nullptr)
{
    assert(bgr.channels() == 3 && bgr.depth == U8);
    mr_context *c = detail::ctx_for(width, height, device);
    Mat out(height, width, U8, 1);
    if (exposure) detail::check(c, mr_ingest_frame_exposure(c, bgr.ptr<uint8_t>(), bgr.cols, bgr.rows, exposure, out.ptr<uint8_t>()));
    else detail::check(c, mr_ingest_frame(c, bgr.ptr<uint8_t>(), bgr.cols, bgr.rows, out.ptr<uint8_t>()));
    return out;
}

}  // namespace mr
