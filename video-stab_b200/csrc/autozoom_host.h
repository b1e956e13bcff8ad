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
//     complement of what a 4-connected flood from outside the image can reach without crossing the contour's border pixels
//     (in place, in the same padded image: the flood only ever touches the free corners).
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

// work: (h+2) x (w+2) signed bytes, 0 / 1 with a zero frame.  Returns the external contours in cv::findContours order and leaves
// the traced border pixels marked (2, or -126 where the pixel to the right is background).
static inline void find_external_contours_padded(signed char* work, int w, int h, std::vector<std::vector<Pt>>& contours) {
    const int step = w + 2;
    std::vector<std::vector<Pt>> found;
    for (int y = 1; y <= h; ++y) {
        signed char* row = work + (size_t)y * step;
        int prev = 0, lnbd_x = 0;
        for (int x = 1; x <= w + 1; ++x) {
            int p = row[x];
            if (p == prev) {
                // runs of equal pixels (almost the whole image) are skipped 32, then eight at a time
                const unsigned long long pat = 0x0101010101010101ull * (unsigned char)prev;
                while (x + 32 <= w + 1) {
                    unsigned long long v[4];
                    std::memcpy(v, row + x, 32);
                    if (((v[0] ^ pat) | (v[1] ^ pat) | (v[2] ^ pat) | (v[3] ^ pat)) != 0) break;
                    x += 32;
                }
                while (x + 8 <= w + 1) {
                    unsigned long long v;
                    std::memcpy(&v, row + x, 8);
                    if (v != pat) break;
                    x += 8;
                }
                if (x > w + 1) break;
                p = row[x];
                if (p == prev) continue;
            }
            if (prev == 0 && p == 1 && !(row[lnbd_x] > 0)) {      // start of an outer border that no traced border encloses
                found.emplace_back();
                fetch_contour(work, step, x, y, found.back());
            }
            prev = row[x];                                        // (possibly just marked)
            if (prev & -2) lnbd_x = x;
        }
    }
    contours.assign(found.rbegin(), found.rend());                // last found first
}
static inline void pad_mask(const uint8_t* mask, int w, int h, size_t stride, std::vector<signed char>& work) {
    const int step = w + 2;
    work.resize((size_t)step * (h + 2));                // every byte is written below: the zero frame, then the rows
    std::memset(work.data(), 0, step);
    std::memset(work.data() + (size_t)(h + 1) * step, 0, step);
    for (int y = 0; y < h; ++y) {
        const uint8_t* m = mask + (size_t)y * stride;
        signed char* d = work.data() + (size_t)(y + 1) * step + 1;
        d[-1] = 0; d[w] = 0;
        int x = 0;
        for (; x + 8 <= w; x += 8) {                    // eight pixels at a time: any non-zero byte -> 1
            unsigned long long v;
            std::memcpy(&v, m + x, 8);
            v = (v & 0x0F0F0F0F0F0F0F0Full) | ((v >> 4) & 0x0F0F0F0F0F0F0F0Full);
            v |= (v >> 2) & 0x0303030303030303ull;
            v |= (v >> 1);
            v &= 0x0101010101010101ull;
            std::memcpy(d + x, &v, 8);
        }
        for (; x < w; ++x) d[x] = m[x] ? 1 : 0;
    }
}
static inline void find_external_contours(const uint8_t* mask, int w, int h, size_t stride, std::vector<std::vector<Pt>>& contours,
                                          std::vector<signed char>& work) {
    pad_mask(mask, w, h, stride, work);
    find_external_contours_padded(work.data(), w, h, contours);
}

// AutoZoomCrop.cpp:9-84.  `work` doubles as the filled contour mask: a pixel is OUTSIDE the drawn contour iff it is marked EXT.
// The reference recounts the zeros on all four border lines of the rectangle in every iteration of its shrinking loop; the
// rectangle only ever shrinks, so the same four counts are maintained incrementally here (a side is recounted when it moves, the
// two perpendicular sides give up the pixels they lose) - identical counts, a fraction of the pixel visits.
enum : signed char { AZ_BORDER = 3, AZ_EXT = 4 };
struct BorderCounts {
    const signed char* work;
    int step;
    bool have = false;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;                   // inclusive corners of the last rectangle (x1 = x0 + width - 1 ...)
    unsigned cTop = 0, cBottom = 0, cLeft = 0, cRight = 0;
    bool zero(int y, int x) const { return work[(size_t)(y + 1) * step + x + 1] == AZ_EXT; }
    unsigned row(int y, int xa, int xb) const { unsigned c = 0; for (int x = xa; x <= xb; ++x) c += zero(y, x); return c; }
    unsigned col(int x, int ya, int yb) const { unsigned c = 0; for (int y = ya; y <= yb; ++y) c += zero(y, x); return c; }
    void update(const Rect& bb) {
        const int nx0 = bb.x, ny0 = bb.y, nx1 = bb.x + bb.width - 1, ny1 = bb.y + bb.height - 1;
        const bool shrink = have && nx0 >= x0 && ny0 >= y0 && nx1 <= x1 && ny1 <= y1 && nx0 <= nx1 && ny0 <= ny1;
        if (!shrink) {
            cTop = row(ny0, nx0, nx1); cBottom = row(ny1, nx0, nx1); cLeft = col(nx0, ny0, ny1); cRight = col(nx1, ny0, ny1);
        } else {
            if (ny0 != y0) cTop = row(ny0, nx0, nx1); else cTop -= row(y0, x0, nx0 - 1) + row(y0, nx1 + 1, x1);
            if (ny1 != y1) cBottom = row(ny1, nx0, nx1); else cBottom -= row(y1, x0, nx0 - 1) + row(y1, nx1 + 1, x1);
            if (nx0 != x0) cLeft = col(nx0, ny0, ny1); else cLeft -= col(x0, y0, ny0 - 1) + col(x0, ny1 + 1, y1);
            if (nx1 != x1) cRight = col(nx1, ny0, ny1); else cRight -= col(x1, y0, ny0 - 1) + col(x1, ny1 + 1, y1);
        }
        have = true; x0 = nx0; y0 = ny0; x1 = nx1; y1 = ny1;
    }
};
static inline bool check_interior_exterior(BorderCounts& bc, const Rect& bb, int& top, int& bottom, int& left, int& right) {
    if (bb.width <= 0 || bb.height <= 0) {
        // degenerate rectangle: the reference's four loops read (at most) one line of an empty ROI; count it the plain way
        bc.have = false;
        bc.cTop = bc.cBottom = bc.cLeft = bc.cRight = 0;
        for (int x = 0; x < bb.width; ++x) { bc.cTop += bc.zero(bb.y, bb.x + x); bc.cBottom += bc.zero(bb.y + bb.height - 1, bb.x + x); }
        for (int y = 0; y < bb.height; ++y) { bc.cLeft += bc.zero(bb.y + y, bb.x); bc.cRight += bc.zero(bb.y + y, bb.x + bb.width - 1); }
    } else bc.update(bb);
    const unsigned cTop = bc.cTop, cBottom = bc.cBottom, cLeft = bc.cLeft, cRight = bc.cRight;
    const bool ok = !(cTop | cBottom | cLeft | cRight);
    if (cTop > cBottom) {
        if (cTop > cLeft && cTop > cRight) top = 1;
    } else if (cBottom > cLeft && cBottom > cRight) bottom = 1;
    if (cLeft >= cRight) {
        if (cLeft >= cBottom && cLeft >= cTop) left = 1;
    } else if (cRight >= cTop && cRight >= cBottom) right = 1;
    return ok;
}

// The crop rectangle of AutoZoomCrop.cpp:141-223 from the padded 0/1 content image (modified in place).  Returns false when
// there is no contour (the reference then returns the frame unchanged).  An empty rectangle also means "return the frame".
static inline bool crop_rect_from_padded(signed char* work, int w, int h, Rect* out) {
    std::vector<std::vector<Pt>> contours;
    find_external_contours_padded(work, w, h, contours);
    if (contours.empty()) return false;
    size_t id = 0, max_size = 0;
    for (size_t i = 0; i < contours.size(); ++i)
        if (contours[i].size() > max_size) { max_size = contours[i].size(); id = i; }
    const std::vector<Pt>& c = contours[id];
    const int step = w + 2;
    // ---- cv::drawContours(contourMask, contours, id, 255, FILLED) = this contour's border pixels plus everything they enclose.
    //      Its border: consecutive vertices are joined by horizontal, vertical or 45-degree runs of border pixels.
    for (size_t i = 0; i < c.size(); ++i) {
        const Pt a = c[i], b = c[(i + 1) % c.size()];
        const int sx = (b.x > a.x) - (b.x < a.x), sy = (b.y > a.y) - (b.y < a.y);
        int x = a.x, y = a.y;
        for (;;) {
            work[(size_t)(y + 1) * step + x + 1] = AZ_BORDER;
            if (x == b.x && y == b.y) break;
            x += sx; y += sy;
        }
    }
    //      Everything else: a 4-connected flood from the frame that never enters a border pixel marks the OUTSIDE (it only
    //      ever visits the corners the content leaves free, other components included: they are not part of the drawn contour).
    //      Span by span: a seed is widened to the run of free pixels around it, the rows above and below are searched for the
    //      starts of free runs under that span (every pixel is marked once and looked at a few times, no per-pixel stack traffic).
    {
        auto is_free = [&](int q) { const signed char v = work[q]; return v != AZ_BORDER && v != AZ_EXT; };
        std::vector<int> stack;
        stack.push_back(0);
        while (!stack.empty()) {
            const int p = stack.back();
            stack.pop_back();
            if (!is_free(p)) continue;
            const int y = p / step, row = y * step;
            int xl = p - row, xr = xl;
            while (xl > 0 && is_free(row + xl - 1)) --xl;
            while (xr < w + 1 && is_free(row + xr + 1)) ++xr;
            std::memset(work + row + xl, AZ_EXT, (size_t)(xr - xl + 1));
            for (int dy = -1; dy <= 1; dy += 2) {
                const int ny = y + dy;
                if (ny < 0 || ny > h + 1) continue;
                const int nrow = ny * step;
                bool in_run = false;
                for (int x = xl; x <= xr; ++x) {
                    const bool f = is_free(nrow + x);
                    if (f && !in_run) stack.push_back(nrow + x);
                    in_run = f;
                }
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
    BorderCounts bc{work, step};
    while (minX < maxX && minY < maxY) {
        bb.x = xs[minX]; bb.y = ys[minY]; bb.width = xs[maxX] - xs[minX]; bb.height = ys[maxY] - ys[minY];
        int t = 0, b = 0, l = 0, r = 0;
        if (check_interior_exterior(bc, bb, t, b, l, r)) break;
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
    const int x1 = std::max(bb.x, 0), y1 = std::max(bb.y, 0), x2 = std::min(bb.x + bb.width, w), y2 = std::min(bb.y + bb.height, h);
    Rect v;
    if (x2 > x1 && y2 > y1) { v.x = x1; v.y = y1; v.width = x2 - x1; v.height = y2 - y1; }
    *out = v;
    return true;
}
static inline bool crop_rect_from_mask(const uint8_t* mask, int w, int h, size_t stride, Rect* out) {
    static thread_local std::vector<signed char> work;   // kept between calls: a fresh 8 MB vector costs more in page faults than the search
    pad_mask(mask, w, h, stride, work);
    return crop_rect_from_padded(work.data(), w, h, out);
}

}  // namespace azc
