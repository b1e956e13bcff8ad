// autozoom_host.h — the host half of vs::AutoZoomCrop::autoZoomCrop (reference src/AutoZoomCrop.cpp:141-223).
//
// The reference itself runs this part on the CPU: it downloads the content mask and calls cv::findContours / cv::drawContours,
// sorts the contour's vertices and shrinks a rectangle until its four border lines lie inside the contour (AutoZoomCrop.cpp:
// 141-204).  It is an inherently sequential border-following + greedy loop over a few thousand vertices, so it stays on the host
// here too; the image-sized work either side of it (gray, threshold, morphological close; crop + resize) is on the device
// (k_autozoom.cu).  Written from the published algorithms:
//   * cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE): Suzuki & Abe border following as OpenCV implements it (raster scan for
//     0 -> 1 transitions not enclosed by an already traced border; 8-connected clockwise/counter-clockwise neighbour search; border
//     pixels marked 2, or -126 when the pixel to their right is background; a vertex is emitted whenever the step direction
//     changes); contours are returned last-found first;
//   * cv::drawContours(FILLED) of one external contour: its border pixels plus everything they enclose.  Computed as the
//     complement of what a 4-connected flood from outside the image can reach without crossing the contour's border pixels.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace azc {

struct Pt { int x, y; };
struct Rect { int x = 0, y = 0, width = 0, height = 0; };

// 8-neighbour step codes, counter-clockwise from east (y grows downwards): E, NE, N, NW, W, SW, S, SE
static const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
static const int DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

// img: (h+2) x (w+2) signed bytes, 0 / 1, with a zero frame of one pixel; origin (x, y) in frame coordinates.
static inline void fetch_contour(signed char* img, int step, int x, int y, std::vector<Pt>& out) {
    const signed char nbd = 2;
    int deltas[16];
    for (int k = 0; k < 8; ++k) deltas[k] = deltas[k + 8] = DY[k] * step + DX[k];
    signed char* i0 = img + (size_t)y * step + x;
    signed char *i1, *i3, *i4 = nullptr;
    Pt pt{x - 1, y - 1};                                     // image coordinates (frame removed)
    int s = 4, s_end = 4, prev_s;
    do {
        s = (s - 1) & 7;
        i1 = i0 + deltas[s];
    } while (*i1 == 0 && s != s_end);
    if (s == s_end) {                                        // single pixel
        *i0 = (signed char)(nbd | -128);
        out.push_back(pt);
        return;
    }
    i3 = i0;
    prev_s = s ^ 4;
    for (;;) {
        s_end = s;
        while (s < 15) {
            i4 = i3 + deltas[++s];
            if (*i4 != 0) break;
        }
        s &= 7;
        if ((unsigned)(s - 1) < (unsigned)s_end) *i3 = (signed char)(nbd | -128);     // the pixel to the right is background
        else if (*i3 == 1) *i3 = nbd;
        if (s != prev_s) {                                   // CHAIN_APPROX_SIMPLE: a vertex where the direction changes
            out.push_back(pt);
            prev_s = s;
        }
        pt.x += DX[s];
        pt.y += DY[s];
        if (i4 == i0 && i3 == i1) break;
        i3 = i4;
        s = (s + 4) & 7;
    }
}

// mask: h x w, nonzero = foreground.  Returns the external contours in cv::findContours order and leaves the marked image in `work`
// ((h+2) x (w+2): 0 background, 1 untouched foreground, 2 / -126 traced border pixels).
static inline void find_external_contours(const uint8_t* mask, int w, int h, size_t stride, std::vector<std::vector<Pt>>& contours,
                                          std::vector<signed char>& work) {
    const int step = w + 2;
    work.assign((size_t)step * (h + 2), 0);
    for (int y = 0; y < h; ++y) {
        const uint8_t* m = mask + (size_t)y * stride;
        signed char* d = work.data() + (size_t)(y + 1) * step + 1;
        for (int x = 0; x < w; ++x) d[x] = m[x] ? 1 : 0;
    }
    std::vector<std::vector<Pt>> found;
    for (int y = 1; y <= h; ++y) {
        signed char* row = work.data() + (size_t)y * step;
        int prev = 0, lnbd_x = 0;
        for (int x = 1; x <= w + 1; ++x) {
            const int p = row[x];
            if (p == prev) continue;
            if (prev == 0 && p == 1 && !(row[lnbd_x] > 0)) {      // start of an outer border that no traced border encloses
                found.emplace_back();
                fetch_contour(work.data(), step, x, y, found.back());
            }
            prev = row[x];                                        // (possibly just marked)
            if (prev & -2) lnbd_x = x;
        }
    }
    contours.assign(found.rbegin(), found.rend());                // last found first
}

// AutoZoomCrop.cpp:9-84
static inline bool check_interior_exterior(const std::vector<uint8_t>& filled, int w, const Rect& bb, int& top, int& bottom, int& left, int& right) {
    bool ok = true;
    unsigned cTop = 0, cBottom = 0, cLeft = 0, cRight = 0;
    auto at = [&](int y, int x) { return filled[(size_t)(bb.y + y) * w + bb.x + x]; };
    for (int x = 0; x < bb.width; ++x) if (at(0, x) == 0) { ok = false; ++cTop; }
    for (int x = 0; x < bb.width; ++x) if (at(bb.height - 1, x) == 0) { ok = false; ++cBottom; }
    for (int y = 0; y < bb.height; ++y) if (at(y, 0) == 0) { ok = false; ++cLeft; }
    for (int y = 0; y < bb.height; ++y) if (at(y, bb.width - 1) == 0) { ok = false; ++cRight; }
    if (cTop > cBottom) {
        if (cTop > cLeft && cTop > cRight) top = 1;
    } else if (cBottom > cLeft && cBottom > cRight) bottom = 1;
    if (cLeft >= cRight) {
        if (cLeft >= cBottom && cLeft >= cTop) left = 1;
    } else if (cRight >= cTop && cRight >= cBottom) right = 1;
    return ok;
}

// The crop rectangle of AutoZoomCrop.cpp:141-223 from the (closed) content mask.  Returns false when there is no contour (the
// reference then returns the frame unchanged).  An empty rectangle (width or height <= 0) also means "return the frame".
static inline bool crop_rect_from_mask(const uint8_t* mask, int w, int h, size_t stride, Rect* out) {
    std::vector<std::vector<Pt>> contours;
    std::vector<signed char> work;
    find_external_contours(mask, w, h, stride, contours, work);
    if (contours.empty()) return false;
    size_t id = 0, max_size = 0;
    for (size_t i = 0; i < contours.size(); ++i)
        if (contours[i].size() > max_size) { max_size = contours[i].size(); id = i; }
    const std::vector<Pt>& c = contours[id];
    // ---- cv::drawContours(contourMask, contours, id, 255, FILLED): border + enclosed pixels.  Re-trace only this contour on a
    //      clean copy so that `border` holds exactly its border pixels, then flood the outside.
    const int step = w + 2;
    std::vector<uint8_t> border((size_t)step * (h + 2), 0);
    {
        // walk the polygon: consecutive vertices are joined by horizontal, vertical or 45-degree runs of border pixels
        for (size_t i = 0; i < c.size(); ++i) {
            Pt a = c[i], b = c[(i + 1) % c.size()];
            const int sx = (b.x > a.x) - (b.x < a.x), sy = (b.y > a.y) - (b.y < a.y);
            int x = a.x, y = a.y;
            for (;;) {
                border[(size_t)(y + 1) * step + x + 1] = 1;
                if (x == b.x && y == b.y) break;
                x += sx; y += sy;
            }
        }
    }
    std::vector<uint8_t> filled((size_t)w * h, 255);
    {
        // 4-connected flood of everything reachable from the frame without entering a border pixel
        std::vector<uint8_t> seen((size_t)step * (h + 2), 0);
        std::vector<int> stack;
        stack.push_back(0);
        seen[0] = 1;
        while (!stack.empty()) {
            const int p = stack.back();
            stack.pop_back();
            const int y = p / step, x = p - y * step;
            if (x >= 1 && x <= w && y >= 1 && y <= h) filled[(size_t)(y - 1) * w + (x - 1)] = 0;
            const int nx[4] = {x + 1, x - 1, x, x}, ny[4] = {y, y, y + 1, y - 1};
            for (int k = 0; k < 4; ++k) {
                if (nx[k] < 0 || nx[k] > w + 1 || ny[k] < 0 || ny[k] > h + 1) continue;
                const int q = ny[k] * step + nx[k];
                if (seen[q] || border[q]) continue;
                seen[q] = 1;
                stack.push_back(q);
            }
        }
    }
    // ---- AutoZoomCrop.cpp:161-204
    std::vector<int> xs(c.size()), ys(c.size());
    for (size_t i = 0; i < c.size(); ++i) { xs[i] = c[i].x; ys[i] = c[i].y; }
    std::sort(xs.begin(), xs.end());
    std::sort(ys.begin(), ys.end());
    unsigned minX = 0, maxX = (unsigned)(xs.size() - 1), minY = 0, maxY = (unsigned)(ys.size() - 1);
    Rect bb;
    while (minX < maxX && minY < maxY) {
        bb.x = xs[minX]; bb.y = ys[minY]; bb.width = xs[maxX] - xs[minX]; bb.height = ys[maxY] - ys[minY];
        int t = 0, b = 0, l = 0, r = 0;
        // cv::Mat::operator()(Rect) of an empty ROI is still valid; a zero-sized one makes the loops of the check empty
        if (check_interior_exterior(filled, w, bb, t, b, l, r)) break;
        if (l) ++minX;
        if (r) --maxX;
        if (t) ++minY;
        if (b) --maxY;
    }
    // ---- :206-223 aspect ratio, centring, clamping
    const double ar = (double)w / (double)h;
    const int new_w = (int)(bb.height * ar);
    const int cx = bb.x + bb.width / 2;
    bb.width = new_w;
    bb.x = cx - new_w / 2;
    if (bb.x < 0) bb.x = 0;
    if (bb.x + bb.width > w) bb.x = w - bb.width;
    // interiorBB &= Rect(0, 0, cols, rows)
    const int x1 = std::max(bb.x, 0), y1 = std::max(bb.y, 0), x2 = std::min(bb.x + bb.width, w), y2 = std::min(bb.y + bb.height, h);
    Rect v;
    if (x2 > x1 && y2 > y1) { v.x = x1; v.y = y1; v.width = x2 - x1; v.height = y2 - y1; }
    *out = v;
    return true;
}

}  // namespace azc
