// TEST INFRASTRUCTURE ONLY (oracle/): satisfies the 11 undefined cv::* symbols of the reference's shipped binary
// Segmentation/Segmentation/cython/src/liblpbox_solver.so so that it can be dlopen()ed without OpenCV -- and implements just
// enough of them for the binary's OWN image front-end, `LPboxADMMsolver::ADMM_bqp_unconstrained_init` (SEG.cpp:658-810), to run:
//
//   cv::imread(path, 0)         -> reads a raw grey image this harness wrote under that path ("LPBXRAW8" rows cols + pixels)
//   cv::resize(src, dst, Size(), fx, fy) -> copy; only fx == fy == 1 (numNodes == rows * cols) is supported
//   Mat::convertTo(dst, CV_64F) -> uint8 -> double element-wise into the caller's header (cv2eigen, opencv2/core/eigen.hpp)
//   cv::transpose(src, dst)     -> 2-D transpose of doubles (in place for the square case cv2eigen uses)
//   Mat::t()                    -> unsupported (cv2eigen takes that route for NON-square images): use square images
//
// With these the harness (ref_harness.seg_member_*) can drive the binary's member entry points that need the state `_init`
// builds -- ADMM_bqp_unconstrained_legacy / _l2f / get_x_iters_d / get_x_sol / get_final_obj (SEG.cpp:868-1380) -- which is what
// pins the early-fixing compaction of the segmentation path to the reference itself.
// cv::Mat is laid out as in OpenCV 4.4 (modules/core/include/opencv2/core/mat.hpp), the version the binary was built against.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {
struct MatSize { int *p; };
struct MatStep { size_t *p; size_t buf[2]; };
struct Mat {                     // OpenCV 4.4, x86-64: 96 bytes
    int flags, dims, rows, cols;
    unsigned char *data;
    const unsigned char *datastart, *dataend, *datalimit;
    void *allocator, *u;
    MatSize size;
    MatStep step;
};
static_assert(sizeof(Mat) == 96, "cv::Mat layout");
struct IOArray { int flags; void *obj; int sz_w, sz_h; };   // cv::_InputArray / _OutputArray
constexpr int MAGIC = 0x42FF0000, CONT = 1 << 14, DEPTH_MASK = 7, CV_8U = 0, CV_64F = 6;
size_t elem_size(int type) { static const size_t d[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return d[type & DEPTH_MASK] * (size_t)(((type >> 3) & 511) + 1); }
void die(const char *what) { std::fprintf(stderr, "cvstub: %s\n", what); std::abort(); }
void mat_alloc(Mat *m, int rows, int cols, int type) {
    m->flags = MAGIC | CONT | (type & 0xFFF);
    m->dims = 2; m->rows = rows; m->cols = cols;
    const size_t es = elem_size(type);
    m->data = (unsigned char *)std::malloc((size_t)rows * cols * es + 64);
    if (!m->data) die("out of memory");
    m->datastart = m->data; m->dataend = m->datalimit = m->data + (size_t)rows * cols * es;
    m->allocator = nullptr; m->u = nullptr;           // u == NULL: Mat::release() never calls deallocate(); the pixels leak (test harness)
    m->size.p = &m->rows;
    m->step.p = m->step.buf; m->step.buf[0] = (size_t)cols * es; m->step.buf[1] = es;
}
}  // namespace

#define STUB(sym) extern "C" void sym() { std::fprintf(stderr, "cvstub: %s called\n", #sym); std::abort(); }
STUB(_ZN2cv5errorEiRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEPKcS9_i)
STUB(_ZNK2cv3Mat1tEv)
// cv::imwrite: save_img() is outside the path; accept and drop
extern "C" bool _ZN2cv7imwriteERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERKNS_11_InputArrayERKSt6vectorIiSaIiEE() { return true; }
extern "C" void _ZN2cv3Mat10deallocateEv(Mat *) {}
extern "C" void _ZN2cv3Mat20updateContinuityFlagEv(Mat *) {}
extern "C" void _ZN2cv8fastFreeEPv(void *p) { std::free(p); }

// void cv::Mat::create(int ndims, const int* sizes, int type)
extern "C" void _ZN2cv3Mat6createEiPKii(Mat *m, int ndims, const int *sizes, int type) {
    if (ndims != 2) die("Mat::create: only 2-D");
    if (m->data && m->dims == 2 && m->rows == sizes[0] && m->cols == sizes[1] && (m->flags & 0xFFF) == (type & 0xFFF)) return;
    mat_alloc(m, sizes[0], sizes[1], type);
}

// cv::Mat cv::imread(const std::string& filename, int flags)  (returned through the hidden pointer)
extern "C" Mat *_ZN2cv6imreadERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEi(Mat *ret, const std::string *path, int) {
    std::FILE *f = std::fopen(path->c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cvstub: imread cannot open %s\n", path->c_str()); std::abort(); }
    char magic[8]; int32_t rc[2];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "LPBXRAW8", 8) != 0 || std::fread(rc, 4, 2, f) != 2) die("imread: not a raw image written by the harness");
    std::memset(ret, 0, sizeof(Mat));
    mat_alloc(ret, rc[0], rc[1], CV_8U);
    if (std::fread(ret->data, 1, (size_t)rc[0] * rc[1], f) != (size_t)rc[0] * rc[1]) die("imread: short file");
    std::fclose(f);
    return ret;
}

// void cv::resize(InputArray src, OutputArray dst, Size dsize, double fx, double fy, int interpolation)
extern "C" void _ZN2cv6resizeERKNS_11_InputArrayERKNS_12_OutputArrayENS_5Size_IiEEddi(const IOArray *src, const IOArray *dst, uint64_t, double fx, double fy, int) {
    const Mat *s = (const Mat *)src->obj;
    Mat *d = (Mat *)dst->obj;
    const int rows = (int)std::lround(s->rows * fy), cols = (int)std::lround(s->cols * fx);
    if (rows != s->rows || cols != s->cols) die("resize: only scale 1 (numNodes == rows * cols) is supported by the stub");
    mat_alloc(d, rows, cols, s->flags & 0xFFF);
    for (int r = 0; r < rows; ++r) std::memcpy(d->data + (size_t)r * d->step.buf[0], s->data + (size_t)r * s->step.p[0], (size_t)cols * elem_size(s->flags));
}

// void cv::Mat::convertTo(OutputArray m, int rtype, double alpha, double beta) const
extern "C" void _ZNK2cv3Mat9convertToERKNS_12_OutputArrayEidd(const Mat *self, const IOArray *out, int rtype, double alpha, double beta) {
    Mat *d = (Mat *)out->obj;
    if ((self->flags & DEPTH_MASK) != CV_8U || (rtype & DEPTH_MASK) != CV_64F || alpha != 1.0 || beta != 0.0) die("convertTo: only uint8 -> double");
    if (!d->data || d->rows != self->rows || d->cols != self->cols || (d->flags & DEPTH_MASK) != CV_64F) die("convertTo: destination header mismatch (non-square image?)");
    for (int r = 0; r < self->rows; ++r) {
        const unsigned char *sp = self->data + (size_t)r * self->step.p[0];
        double *dp = (double *)(d->data + (size_t)r * d->step.p[0]);
        for (int c = 0; c < self->cols; ++c) dp[c] = (double)sp[c];
    }
}

// void cv::transpose(InputArray src, OutputArray dst)
extern "C" void _ZN2cv9transposeERKNS_11_InputArrayERKNS_12_OutputArrayE(const IOArray *src, const IOArray *dst) {
    const Mat *s = (const Mat *)src->obj;
    Mat *d = (Mat *)dst->obj;
    if ((s->flags & DEPTH_MASK) != CV_64F) die("transpose: only double");
    if (s->data == d->data) {
        if (s->rows != s->cols) die("transpose: in place needs a square matrix");
        for (int r = 0; r < s->rows; ++r)
            for (int c = r + 1; c < s->cols; ++c) {
                double *a = (double *)(d->data + (size_t)r * d->step.p[0]) + c, *b = (double *)(d->data + (size_t)c * d->step.p[0]) + r;
                const double t = *a; *a = *b; *b = t;
            }
        return;
    }
    if (!d->data || d->rows != s->cols || d->cols != s->rows) die("transpose: destination header mismatch");
    for (int r = 0; r < s->rows; ++r)
        for (int c = 0; c < s->cols; ++c)
            ((double *)(d->data + (size_t)c * d->step.p[0]))[r] = ((const double *)(s->data + (size_t)r * s->step.p[0]))[c];
}
