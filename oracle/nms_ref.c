/* oracle/nms_ref.c -- TEST INFRASTRUCTURE (checker only; never linked into the product).
 *
 * CPU restatement of the rotated-BEV IoU / greedy NMS that the reference calls through its binary-only extension
 *   unidistill/layers/head/det3d/generate_proposals/iou3d_nms_cuda.cpython-36m-x86_64-linux-gnu.so
 * (exports nms_gpu / boxes_iou_bev_gpu / boxes_overlap_bev_gpu / nms_normal_gpu; call site
 *  generate_proposals/centerpoint_gen_proposals.py:85-106).  No source is shipped with the reference; the symbol
 * names and the (N,7) [x,y,z,dx,dy,dz,heading] box contract identify it as OpenPCDet's `iou3d_nms` op (v0.5 layout),
 * whose published algorithm is restated here in single precision:
 *   - both rectangles' corners rotated about their centres,
 *   - intersection points of every edge pair (bounding-box pre-test, then the signed-area straddle test),
 *   - corners of one box inside the other with a 1e-2 margin,
 *   - the points sorted by polar angle about their mean (bubble sort, strict >), fan area, |area| / 2,
 *   - IoU = overlap / max(area_a + area_b - overlap, 1e-8),
 *   - NMS: boxes arrive sorted by descending score; box i is kept unless an earlier KEPT box j has IoU(j, i) > thresh.
 * PINNED AGAINST THE REFERENCE BINARY: the extension cannot be imported (cp36 / torch-1.9 ABI, sm_61-86 cubins), but
 * its CPU routine boxes_iou_bev_cpu can be executed in the build container through oracle/refshim/iou3d_shim.cpp;
 * tests/golden/iou3d_ref.npz holds its output for 78 000 box pairs (random + degenerate cases) and
 * ud_ref_boxes_iou_bev reproduces every value bit for bit (tests/test_oracle_nms.py).  The GPU routines nms_gpu /
 * boxes_iou_bev_gpu of the binary cannot run here (no cubin for sm_100); the sweep restated in ud_ref_nms_bev is the
 * published one.  Further checks: known answers, exact float64 polygon clipping.  What the binary itself shows
 * (objdump / byte search, no execution): its .rodata holds the float constants 1e-8 and 1e-2 side by side (EPS and the
 * in-box MARGIN of this algorithm), and boxes_iou_bev_cpu calls sincosf / sinf, atan2f and fmaxf -- the corner
 * rotation, the polar-angle comparison and the max(union, EPS) of the restatement below.  The sm_86 SASS of its
 * nms_kernel (cuobjdump -sass) compares with FSETP.GT against the threshold parameter (strict >) and selects the
 * first column of a diagonal 64x64 block by ISETP.NE on (blockIdx.y, blockIdx.x) (start = threadIdx.x + 1 there, else 0):
 * the suppression rule restated in ud_ref_nms_bev.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define UD_NMS_EPS 1e-8f
#define UD_NMS_MARGIN 1e-2f

typedef struct { float x, y; } pt;

static float cross2(pt a, pt b) { return a.x * b.y - a.y * b.x; }
static float cross3(pt p1, pt p2, pt p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }

static int rect_cross(pt p1, pt p2, pt q1, pt q2) {
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

static int in_box(const float* box, pt p) {
    const float c = cosf(-box[6]), s = sinf(-box[6]);
    const float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    const float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + UD_NMS_MARGIN && fabsf(ry) < box[4] / 2 + UD_NMS_MARGIN;
}

static int seg_intersection(pt p1, pt p0, pt q1, pt q0, pt* ans) {
    if (!rect_cross(p0, p1, q0, q1)) return 0;
    const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    const float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > UD_NMS_EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static void corners(const float* box, pt* out /* 5 */) {
    const float hx = box[3] / 2, hy = box[4] / 2;
    const float c = cosf(box[6]), s = sinf(box[6]);
    const float lx[4] = {-hx, hx, hx, -hx}, ly[4] = {-hy, -hy, hy, hy};
    for (int k = 0; k < 4; ++k) {
        /* rotate (x1|x2, y1|y2) around the centre, written as the reference does: offset from the centre first */
        const float px = (box[0] + lx[k]) - box[0], py = (box[1] + ly[k]) - box[1];
        out[k].x = px * c + py * (-s) + box[0];
        out[k].y = px * s + py * c + box[1];
    }
    out[4] = out[0];
}

float ud_ref_box_overlap(const float* a, const float* b) {
    pt ca[5], cb[5], pts[16], ctr = {0.f, 0.f};
    corners(a, ca);
    corners(b, cb);
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
                ctr.x += pts[cnt].x;
                ctr.y += pts[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (in_box(a, cb[k])) { ctr.x += cb[k].x; ctr.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (in_box(b, ca[k])) { ctr.x += ca[k].x; ctr.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    if (cnt == 0) return 0.f;
    ctr.x /= cnt;
    ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
                pt t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        pt u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
        area += cross2(u, v);
    }
    return fabsf(area) / 2.0f;
}

float ud_ref_iou_bev(const float* a, const float* b) {
    const float sa = a[3] * a[4], sb = b[3] * b[4];
    const float so = ud_ref_box_overlap(a, b);
    return so / fmaxf(sa + sb - so, UD_NMS_EPS);
}

/* boxes_iou_bev_gpu: ans (na, nb) */
void ud_ref_boxes_iou_bev(const float* a, int na, const float* b, int nb, float* ans) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) ans[(size_t)i * nb + j] = ud_ref_iou_bev(a + 7 * i, b + 7 * j);
}

/* nms_gpu: boxes (n,7) sorted by descending score -> keep (indices into boxes, ascending), returns their count */
int ud_ref_nms_bev(const float* boxes, int n, float thresh, int64_t* keep) {
    unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
    int num = 0;
    for (int i = 0; i < n; ++i) {
        if (removed[i]) continue;
        keep[num++] = i;
        for (int j = i + 1; j < n; ++j)
            if (!removed[j] && ud_ref_iou_bev(boxes + 7 * i, boxes + 7 * j) > thresh) removed[j] = 1;
    }
    free(removed);
    return num;
}
