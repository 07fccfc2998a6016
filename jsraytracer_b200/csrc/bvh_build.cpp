// Native restatement of the reference BVH build (src/aggregates.js:65-185:
// BVHAggregateNode.build / split_objects — 8-bin SAH on centre/half-size f32
// boxes, median-of-centres fallback, halving fallback) for meshes too large
// for the Python mirror (jsraytracer_b200/world.py).  Same arithmetic: AABB
// vectors are f32 (`Vec`), every scalar expression is f64.  The tree topology
// decides hit-ID ties and the roofline's node counts, so this follows the
// reference statement by statement instead of using a better builder.
// Setup code, host only.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/jsrt.h"

namespace {

const float FINF = std::numeric_limits<float>::infinity();

struct Box {                      // AABB: center, half_size, min, max (src/geometry.js:77-84)
    float c[3], h[3], mn[3], mx[3];
};

// AABB.fromMinMax src/geometry.js:94-105
Box fromMinMax(const float mn[3], const float mx[3]) {
    Box b;
    for (int i = 0; i < 3; ++i) {
        b.mn[i] = mn[i]; b.mx[i] = mx[i];
        b.c[i] = (float)((1 - 0.5) * (double)mn[i] + 0.5 * (double)mx[i]);     // min.mix(max, 0.5)
        const float d = (float)((double)mx[i] - (double)mn[i]);                  // max.minus(min)
        b.h[i] = (float)((double)d * 0.5);                                       // .times(0.5)
        if (!std::isfinite(mn[i]) && !std::isfinite(mx[i])) {
            if (mn[i] == mx[i]) b.h[i] = 0;
            if (mn[i] == -FINF && mx[i] == FINF) b.c[i] = 0;
        }
    }
    return b;
}
Box emptyBox() {                  // AABB.empty src/geometry.js:91-93
    Box b; for (int i = 0; i < 3; ++i) { b.c[i] = 0; b.h[i] = 0; b.mn[i] = FINF; b.mx[i] = -FINF; } return b;
}
struct Hull {                     // AABB.hull src/geometry.js:121-135, incremental form
    float mn[3] = {FINF, FINF, FINF}, mx[3] = {-FINF, -FINF, -FINF};
    int count = 0;
    void add(const Box& b) { for (int i = 0; i < 3; ++i) { if (b.mn[i] < mn[i]) mn[i] = b.mn[i]; if (b.mx[i] > mx[i]) mx[i] = b.mx[i]; } ++count; }
    Box box() const { return count ? fromMinMax(mn, mx) : emptyBox(); }
};
double surfaceArea(const Box& b) {   // src/geometry.js:160-164
    return 4 * ((double)b.h[0] * b.h[1] + (double)b.h[0] * b.h[2] + (double)b.h[1] * b.h[2]);
}

// ---- median / quickselect with JS array semantics (src/math.js:95-157) --------
// NaN stands for `undefined`: out-of-range reads give it, out-of-range writes
// grow the array, and it compares false against everything.
double aget(const std::vector<double>& a, long i) { return (i >= 0 && i < (long)a.size()) ? a[(size_t)i] : std::nan(""); }
void aset(std::vector<double>& a, long i, double x) { while ((long)a.size() <= i) a.push_back(std::nan("")); a[(size_t)i] = x; }
void aswap(std::vector<double>& a, long i, long j) { double t = aget(a, i); aset(a, i, aget(a, j)); aset(a, j, t); }
int cmp(double a, double b) { return a < b ? -1 : (a > b ? 1 : 0); }

void quickSelectStep(std::vector<double>& arr, long k, long left, long right) {
    while (right > left) {
        if (right - left > 600) {
            const double n = (double)(right - left + 1), m = (double)(k - left + 1), z = std::log(n), s = 0.5 * std::exp(2 * z / 3);
            const double sd = 0.5 * std::sqrt(z * s * (n - s) / n) * (m - n / 2 < 0 ? -1 : 1);
            const long newLeft = std::max(left, (long)std::floor(k - m * s / n + sd));
            const long newRight = std::min(right, (long)std::floor(k + (n - m) * s / n + sd));
            quickSelectStep(arr, k, newLeft, newRight);
        }
        const double t = aget(arr, k);
        long i = left, j = right;
        aswap(arr, left, k);
        if (cmp(aget(arr, right), t) > 0) aswap(arr, left, right);
        while (i < j) {
            aswap(arr, i, j); i++; j--;
            while (cmp(aget(arr, i), t) < 0) i++;
            while (cmp(aget(arr, j), t) > 0) j--;
        }
        if (cmp(aget(arr, left), t) == 0) aswap(arr, left, j);
        else { j++; aswap(arr, j, right); }
        if (j <= k) left = j + 1;
        if (k <= j) right = j - 1;
    }
}
double quickSelect(std::vector<double>& arr, long k) { quickSelectStep(arr, k, 0, (long)arr.size() - 1); return aget(arr, k); }
double median(std::vector<double>& arr) {
    if (arr.empty()) return std::nan("");
    const long len2 = (long)arr.size() / 2;
    if (arr.size() % 2 == 1) return quickSelect(arr, len2);
    const double a = quickSelect(arr, len2);
    const double b = quickSelect(arr, len2 + 1);
    return (a + b) / 2;
}

struct Builder {
    const std::vector<Box>& boxes;
    double maxDepth; int minNodeSize;
    std::vector<jsrt_bvh_node> nodes;
    std::vector<int32_t> leaf_objs;

    int makeLeaf(const std::vector<int>& objs, int depth) {
        Hull h; for (int o : objs) h.add(boxes[o]);
        return pushNode(depth, true, h.box(), objs);
    }
    int pushNode(int depth, bool leaf, const Box& b, const std::vector<int>& objs) {
        jsrt_bvh_node n{};
        n.depth = depth; n.is_leaf = leaf ? 1 : 0; n.lesser = n.greater = -1;
        n.obj_first = (int)leaf_objs.size(); n.obj_count = leaf ? (int)objs.size() : 0;
        if (leaf) leaf_objs.insert(leaf_objs.end(), objs.begin(), objs.end());
        for (int i = 0; i < 3; ++i) { n.center[i] = b.c[i]; n.half_size[i] = b.h[i]; n.min[i] = b.mn[i]; n.max[i] = b.mx[i]; }
        n.center[3] = 1; n.half_size[3] = 0; n.min[3] = 1; n.max[3] = 1;
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }

    // BVHAggregateNode.split_objects src/aggregates.js:87-185
    bool split(const std::vector<int>& objects, Box& bounds, std::vector<int>& lesser, std::vector<int>& greater) {
        const int binsPerAxis = 8;
        if (objects.size() < 2) return false;
        Hull hb; for (int o : objects) hb.add(boxes[o]);
        bounds = hb.box();
        int best_axis = -1; double best_sep_value = INFINITY, best_cost = INFINITY;
        const double bsa = surfaceArea(bounds);
        for (int axis = 0; axis < 3; ++axis) {
            if (bounds.h[axis] < 0.000001) continue;
            int counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            Hull bb[8];
            const double bmin = bounds.mn[axis], ext = 2 * (double)bounds.h[axis];
            for (int o : objects) {
                double bi = std::floor(binsPerAxis * (((double)boxes[o].c[axis] - bmin) / ext));
                if (bi == binsPerAxis) bi = binsPerAxis - 1;
                if (!(bi >= 0 && bi < binsPerAxis)) bi = bi < 0 ? 0 : binsPerAxis - 1;   // would throw in the reference
                counts[(int)bi]++; bb[(int)bi].add(boxes[o]);
            }
            Box bbox[8]; for (int i = 0; i < 8; ++i) bbox[i] = bb[i].box();
            for (int i = 0; i < binsPerAxis - 1; ++i) {
                Hull h0, h1; int count0 = 0, count1 = 0;
                // hull of hulls: AABB.hull([b0, bins[j].bounds]) recomputes centre/half each
                // time, but only min/max feed the next hull, so accumulating min/max is identical.
                for (int j = 0; j <= i; ++j) if (counts[j] > 0) { h0.add(bbox[j]); count0 += counts[j]; }
                for (int j = i + 1; j < binsPerAxis; ++j) if (counts[j] > 0) { h1.add(bbox[j]); count1 += counts[j]; }
                const double cost = .125 + (count0 * surfaceArea(h0.box()) + count1 * surfaceArea(h1.box())) / bsa;
                if (cost < best_cost && count0 > 0 && count1 > 0) {
                    best_axis = axis;
                    best_sep_value = bmin + ((i + 1) / (double)binsPerAxis) * ext;
                    best_cost = cost;
                }
            }
        }
        if (best_axis < 0) {
            for (int axis = 0; axis < 3; ++axis) {
                std::vector<double> centers; centers.reserve(objects.size());
                for (int o : objects) centers.push_back((double)boxes[o].c[axis]);
                const double med = median(centers);
                Hull h0, h1;
                for (int o : objects) { if ((double)boxes[o].c[axis] < med) h0.add(boxes[o]); }
                for (int o : objects) { if ((double)boxes[o].c[axis] >= med) h1.add(boxes[o]); }
                const double cost = .125 + (h0.count * surfaceArea(h0.box()) + h1.count * surfaceArea(h1.box())) / bsa;
                if (cost < best_cost && h0.count > 0 && h1.count > 0) { best_axis = axis; best_sep_value = med; best_cost = cost; }
            }
            if (best_axis < 0) {
                const size_t sp = objects.size() / 2;
                lesser.assign(objects.begin(), objects.begin() + sp);
                greater.assign(objects.begin() + sp, objects.end());
                return true;
            }
        }
        for (int o : objects) { if ((double)boxes[o].c[best_axis] < best_sep_value) lesser.push_back(o); }
        for (int o : objects) { if ((double)boxes[o].c[best_axis] >= best_sep_value) greater.push_back(o); }
        return true;
    }

    // BVHAggregateNode.build src/aggregates.js:65-86
    int build(const std::vector<int>& objects, int depth) {
        if ((double)depth >= maxDepth || (int)objects.size() <= minNodeSize) return makeLeaf(objects, depth);
        Box bounds; std::vector<int> lesser, greater;
        if (!split(objects, bounds, lesser, greater)) return makeLeaf(objects, depth);
        const int me = pushNode(depth, false, bounds, {});
        const int l = build(lesser, depth + 1);
        const int g = build(greater, depth + 1);
        nodes[me].lesser = l; nodes[me].greater = g;
        return me;
    }
};

struct BvhHandle { std::vector<jsrt_bvh_node> nodes; std::vector<int32_t> leaf_objs; };

}  // namespace

extern "C" {

jsrt_bvh* jsrt_bvh_build(int n, const float* center, const float* half_size, const float* bmin, const float* bmax,
                         double max_depth, int min_node_size) {
    std::vector<Box> boxes((size_t)n);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) { boxes[i].c[k] = center[3 * i + k]; boxes[i].h[k] = half_size[3 * i + k]; boxes[i].mn[k] = bmin[3 * i + k]; boxes[i].mx[k] = bmax[3 * i + k]; }
    Builder b{boxes, max_depth, min_node_size, {}, {}};
    std::vector<int> all((size_t)n); for (int i = 0; i < n; ++i) all[i] = i;
    b.build(all, 0);
    auto* h = new BvhHandle{std::move(b.nodes), std::move(b.leaf_objs)};
    return (jsrt_bvh*)h;
}
int jsrt_bvh_node_count(const jsrt_bvh* h) { return (int)((const BvhHandle*)h)->nodes.size(); }
int jsrt_bvh_leaf_object_count(const jsrt_bvh* h) { return (int)((const BvhHandle*)h)->leaf_objs.size(); }
int jsrt_bvh_copy(const jsrt_bvh* h, jsrt_bvh_node* nodes, int32_t* leaf_objects) {
    const BvhHandle* b = (const BvhHandle*)h;
    if (nodes) memcpy(nodes, b->nodes.data(), b->nodes.size() * sizeof(jsrt_bvh_node));
    if (leaf_objects) memcpy(leaf_objects, b->leaf_objs.data(), b->leaf_objs.size() * sizeof(int32_t));
    return 0;
}
void jsrt_bvh_free(jsrt_bvh* h) { delete (BvhHandle*)h; }

}  // extern "C"
