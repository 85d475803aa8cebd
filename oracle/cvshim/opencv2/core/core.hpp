// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the parts of OpenCV's cv::Mat API that
// /root/reference/ProblemSets/ps2_cpp/lib/DisparitySSD.cpp and lib/DisparityNCorr.cpp touch, so that
// the reference's own translation units can be compiled *in place, unmodified* into oracle/_ref/
// (OpenCV C++ is not installed in this image).  Nothing under oracle/ is linked into, or called by, the product.
//
// Behavioural notes that matter for parity:
//  * Mat::at<T>(r, c) is unchecked, exactly like release-mode OpenCV: data + r*step + c*sizeof(T).
//    The reference reads up to windowRad elements before/after a padded row; inside the buffer that
//    wraps into the neighbouring row, and for the first/last row it leaves the allocation.
//  * create() puts a zero-filled guard band on both sides of the pixel buffer so that those
//    out-of-allocation reads are deterministic (0.0f) instead of heap garbage.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#define CV_8U 0
#define CV_8S 1
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };

struct Point { int x = 0, y = 0; Point() = default; Point(int x_, int y_) : x(x_), y(y_) {} };
// Like cv::Rect_<int>: any integral argument mix narrows to int (DisparityNCorr.cpp:46,52 pass size_t).
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    template <typename A, typename B, typename C, typename D>
    Rect(A x_, B y_, C w_, D h_) : x(int(x_)), y(int(y_)), width(int(w_)), height(int(h_)) {}
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    // Wrap caller memory (no ownership) — used by the C driver for the inputs.
    Mat(int r, int c, int type, void* ext, size_t stepBytes)
        : rows(r), cols(c), step(stepBytes), data(static_cast<unsigned char*>(ext)), _type(type) {}

    static size_t elemSizeOf(int type) {
        switch (type & 7) {
        case CV_8U: case CV_8S: return 1;
        case CV_32F: return 4;
        default: return 1;
        }
    }
    int type() const { return _type; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t elemSize() const { return elemSizeOf(_type); }

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == _type) return;
        const size_t es = elemSizeOf(type);
        const size_t bytes = size_t(r) * size_t(c) * es;
        _buf.reset(new unsigned char[bytes + 2 * GUARD], std::default_delete<unsigned char[]>());
        std::memset(_buf.get(), 0, bytes + 2 * GUARD);
        data = _buf.get() + GUARD;
        rows = r; cols = c; _type = type; step = size_t(c) * es;
    }

    template <typename T> T& at(int r, int c) {
        return *reinterpret_cast<T*>(data + ptrdiff_t(r) * ptrdiff_t(step) + ptrdiff_t(c) * ptrdiff_t(sizeof(T)));
    }
    template <typename T> const T& at(int r, int c) const {
        return *reinterpret_cast<const T*>(data + ptrdiff_t(r) * ptrdiff_t(step) + ptrdiff_t(c) * ptrdiff_t(sizeof(T)));
    }

    // Region of interest: a header onto the same buffer (DisparityNCorr.cpp:47,53).  OpenCV asserts that the
    // rectangle lies inside the matrix and throws otherwise; the shim does the same.
    Mat operator()(const Rect& r) const {
        if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > cols || r.y + r.height > rows)
            throw std::out_of_range("cv::Mat::operator()(Rect): roi outside the matrix");
        Mat m;
        m.rows = r.height; m.cols = r.width; m.step = step; m._type = _type; m._buf = _buf;
        m.data = data + size_t(r.y) * step + size_t(r.x) * elemSize();
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, _type);
        const size_t es = elemSize();
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + size_t(r) * m.step, data + size_t(r) * step, size_t(cols) * es);
        return m;
    }

    static constexpr size_t GUARD = 1 << 16;   // zero bytes either side of the pixel buffer

private:
    int _type = 0;
    std::shared_ptr<unsigned char> _buf;
};

// BORDER_REPLICATE only (the single mode the ps2 path uses).
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
    (void)borderType;
    const size_t es = src.elemSize();
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
    for (int r = 0; r < out.rows; ++r) {
        int sr = r - top; if (sr < 0) sr = 0; if (sr > src.rows - 1) sr = src.rows - 1;
        for (int c = 0; c < out.cols; ++c) {
            int sc = c - left; if (sc < 0) sc = 0; if (sc > src.cols - 1) sc = src.cols - 1;
            std::memcpy(out.data + size_t(r) * out.step + size_t(c) * es,
                        src.data + size_t(sr) * src.step + size_t(sc) * es, es);
        }
    }
    dst = out;
}

} // namespace cv
