// AMR kd-tree: brick decomposition and view-order traversal, host side (no CUDA).
//
// What it reproduces (node for node, so that the brick list and its order equal the reference's):
//   * Tree.build / Node.add_grids / insert_grids / split_grids / kdtree_get_choices / divide
//       -- yt/utilities/amr_kdtree/amr_kdtree.py:96-118, yt/utilities/lib/amr_kdtools.pyx:136-458,756-823
//   * kd_traverse / depth_traverse / viewpoint_traverse (far child first)
//       -- amr_kdtools.pyx:102-111,565-622,674-755
//   * rank ownership of subtrees (should_i_build / should_i_split / geo_split)
//       -- amr_kdtools.pyx:270-274,460-507,670-676
// How: an index-based node pool (no Python objects, no generators); a traversal writes the whole
// ordered leaf list into caller arrays in one call -- the 20k-leaf Python generator of the
// reference becomes one C call per frame.
#include "../../include/ytb200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct KdNode {
    double le[3], re[3];
    int64_t grid;       // -1: no grid
    int64_t node_id;    // heap numbering: root 1, children 2i, 2i+1
    int left, right, parent;
    int split_dim;      // -1: none
    double split_pos;
};

struct GridBox {
    double le[3], re[3];
    int64_t id;
};

}  // namespace

struct ytb200_kdtree {
    std::vector<KdNode> nodes;
    int rank = 0, size = 1;
    std::string err;

    bool is_leaf(int n) const { return nodes[n].left < 0 || nodes[n].right < 0; }

    // amr_kdtools.pyx:670-676
    bool should_i_build(int n) const
    {
        const int64_t id = nodes[n].node_id;
        if (id < size || id >= 2 * (int64_t)size) return true;
        return id - size == rank;
    }
    // amr_kdtools.pyx:270-274
    bool should_i_split(int n) const { return nodes[n].node_id < size && nodes[n].node_id > 0; }

    int new_node(int parent, const double le[3], const double re[3], int64_t grid, int64_t id)
    {
        KdNode k;
        memcpy(k.le, le, sizeof k.le);
        memcpy(k.re, re, sizeof k.re);
        k.grid = grid; k.node_id = id; k.left = k.right = -1; k.parent = parent;
        k.split_dim = -1; k.split_pos = NAN;
        nodes.push_back(k);
        return (int)nodes.size() - 1;
    }

    // amr_kdtools.pyx:509-534: children inherit the parent's grid
    void divide(int n, int dim, double pos)
    {
        nodes[n].split_dim = dim; nodes[n].split_pos = pos;
        double le[3], re[3];
        memcpy(le, nodes[n].le, sizeof le);
        memcpy(re, nodes[n].re, sizeof re);
        re[dim] = pos;
        const int l = new_node(n, le, re, nodes[n].grid, nodes[n].node_id << 1);
        re[dim] = nodes[n].re[dim];
        le[dim] = pos;
        const int r = new_node(n, le, re, nodes[n].grid, (nodes[n].node_id << 1) + 1);
        nodes[n].left = l; nodes[n].right = r;
    }

    // kdtree_get_choices, amr_kdtools.pyx:756-823
    int get_choices(const std::vector<GridBox> &g, const double lc[3], const double rc[3], double *split,
                    std::vector<char> &less, std::vector<char> &greater) const
    {
        int best_dim = -1;
        size_t my_max = 0;
        std::vector<double> best;
        for (int dim = 0; dim < 3; dim++) {
            std::vector<double> uniq;
            for (const GridBox &b : g) {
                for (int j = 0; j < 2; j++) {
                    const double v = j == 0 ? b.le[dim] : b.re[dim];
                    if (!(lc[dim] < v && v < rc[dim])) continue;
                    if (std::find(uniq.begin(), uniq.end(), v) == uniq.end()) uniq.push_back(v);
                }
            }
            if (uniq.size() > my_max) { best_dim = dim; my_max = uniq.size(); best.swap(uniq); }
        }
        if (best_dim == -1) return -1;
        std::sort(best.begin(), best.end());
        *split = best[(my_max - 1) / 2];
        less.assign(g.size(), 0);
        greater.assign(g.size(), 0);
        for (size_t i = 0; i < g.size(); i++) {
            less[i] = g[i].le[best_dim] < *split;
            greater[i] = g[i].re[best_dim] > *split;
        }
        return best_dim;
    }

    // geo_split, amr_kdtools.pyx:460-507 (halves the GRID's longest axis; both children see the whole grid)
    void geo_split(int n, const GridBox &b)
    {
        int big = 0;
        double my_max = 0.0;
        for (int i = 0; i < 3; i++) {
            const double v = b.re[i] - b.le[i];
            if (v > my_max) { my_max = v; big = i; }
        }
        divide(n, big, (b.re[big] + b.le[big]) / 2.0);
        const int l = nodes[n].left, r = nodes[n].right;
        insert_grid(l, b);
        insert_grid(r, b);
    }

    // insert_grid, amr_kdtools.pyx:164-198
    void insert_grid(int n, const GridBox &b)
    {
        if (!should_i_build(n)) return;
        if (should_i_split(n)) { geo_split(n, b); return; }
        bool contained = true;
        for (int i = 0; i < 3; i++)
            if (b.le[i] > nodes[n].le[i] || b.re[i] < nodes[n].re[i]) contained = false;
        if (contained) { nodes[n].grid = b.id; return; }
        std::vector<GridBox> one(1, b);
        if (split_grids(n, one) == -1) nodes[n].grid = -1;
    }

    // insert_grids, amr_kdtools.pyx:276-316
    void insert_grids(int n, const std::vector<GridBox> &g)
    {
        if (!should_i_build(n) || g.empty()) return;
        if (g.size() == 1) {
            if (should_i_split(n)) { geo_split(n, g[0]); return; }
            bool contained = true;
            for (int i = 0; i < 3; i++) {
                contained = contained && g[0].le[i] <= nodes[n].le[i];
                contained = contained && g[0].re[i] >= nodes[n].re[i];
            }
            if (contained) { nodes[n].grid = g[0].id; return; }
        }
        if (split_grids(n, g) == -1) nodes[n].grid = -1;
    }

    // split_grid / split_grids, amr_kdtools.pyx:318-458
    int split_grids(int n, const std::vector<GridBox> &g)
    {
        if (g.empty()) return 0;
        std::vector<char> less, greater;
        double pos = 0.0;
        const int dim = get_choices(g, nodes[n].le, nodes[n].re, &pos, less, greater);
        if (dim == -1) return -1;
        divide(n, dim, pos);
        std::vector<GridBox> lg, rg;
        for (size_t i = 0; i < g.size(); i++) {
            if (less[i]) lg.push_back(g[i]);
            if (greater[i]) rg.push_back(g[i]);
        }
        const int l = nodes[n].left, r = nodes[n].right;
        if (!lg.empty()) insert_grids(l, lg);
        if (!rg.empty()) insert_grids(r, rg);
        return 0;
    }

    // add_grids, amr_kdtools.pyx:200-268
    void add_grids(int n, const std::vector<GridBox> &g)
    {
        if (!should_i_build(n)) return;
        if (is_leaf(n)) { insert_grids(n, g); return; }
        std::vector<GridBox> lg, rg;
        const int dim = nodes[n].split_dim;
        const double pos = nodes[n].split_pos;
        for (const GridBox &b : g) {
            if (b.le[dim] < pos) lg.push_back(b);
            if (b.re[dim] > pos) rg.push_back(b);
        }
        const int l = nodes[n].left, r = nodes[n].right;
        if (!lg.empty()) add_grids(l, lg);
        if (!rg.empty()) add_grids(r, rg);
    }

    // kd_traverse: leaves with a grid, depth-first left-to-right (viewpoint == nullptr) or with the
    // child farther from the viewpoint first (step_viewpoint: `viewpoint[dim] <= pos` -> right first).
    void traverse(const double *viewpoint, std::vector<int> &out) const
    {
        out.clear();
        if (nodes.empty()) return;
        std::vector<int> stack(1, 0);
        while (!stack.empty()) {
            const int n = stack.back();
            stack.pop_back();
            if (is_leaf(n)) {
                if (nodes[n].grid != -1) out.push_back(n);
                continue;
            }
            int first = nodes[n].left, second = nodes[n].right;
            if (viewpoint && viewpoint[nodes[n].split_dim] <= nodes[n].split_pos) std::swap(first, second);
            stack.push_back(second);
            stack.push_back(first);
        }
    }
};

static thread_local std::string g_kd_err;
extern "C" const char *ytb200_kdtree_last_error(void) { return g_kd_err.c_str(); }

extern "C" int ytb200_kdtree_create(const double left[3], const double right[3], int rank, int size, ytb200_kdtree **out)
{
    if (!left || !right || !out) { g_kd_err = "NULL argument"; return -1; }
    if (size < 1 || rank < 0 || rank >= size) { g_kd_err = "bad rank/size"; return -1; }
    auto *t = new ytb200_kdtree;
    t->rank = rank; t->size = size;
    t->new_node(-1, left, right, -1, 1);
    *out = t;
    return 0;
}

extern "C" int ytb200_kdtree_destroy(ytb200_kdtree *t)
{
    delete t;
    return 0;
}

extern "C" int ytb200_kdtree_add_grids(ytb200_kdtree *t, int64_t n, const double *gles, const double *gres, const int64_t *gids)
{
    if (!t || (n > 0 && (!gles || !gres || !gids))) { g_kd_err = "NULL argument"; return -1; }
    std::vector<GridBox> g((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        for (int j = 0; j < 3; j++) { g[(size_t)i].le[j] = gles[3 * i + j]; g[(size_t)i].re[j] = gres[3 * i + j]; }
        g[(size_t)i].id = gids[i];
    }
    t->add_grids(0, g);
    return 0;
}

extern "C" int64_t ytb200_kdtree_node_count(const ytb200_kdtree *t) { return t ? (int64_t)t->nodes.size() : 0; }

extern "C" int ytb200_kdtree_nodes(const ytb200_kdtree *t, int64_t *node_id, int64_t *grid, int32_t *left, int32_t *right,
                                   int32_t *split_dim, double *split_pos, double *le, double *re)
{
    if (!t) { g_kd_err = "tree is NULL"; return -1; }
    for (size_t i = 0; i < t->nodes.size(); i++) {
        const KdNode &k = t->nodes[i];
        if (node_id) node_id[i] = k.node_id;
        if (grid) grid[i] = k.grid;
        if (left) left[i] = k.left;
        if (right) right[i] = k.right;
        if (split_dim) split_dim[i] = k.split_dim;
        if (split_pos) split_pos[i] = k.split_pos;
        if (le) memcpy(le + 3 * i, k.le, sizeof k.le);
        if (re) memcpy(re + 3 * i, k.re, sizeof k.re);
    }
    return 0;
}

extern "C" int64_t ytb200_kdtree_traverse(const ytb200_kdtree *t, const double *viewpoint, int32_t *out, int64_t cap)
{
    if (!t) { g_kd_err = "tree is NULL"; return -1; }
    std::vector<int> order;
    t->traverse(viewpoint, order);
    if (out) {
        if ((int64_t)order.size() > cap) { g_kd_err = "output too small"; return -1; }
        for (size_t i = 0; i < order.size(); i++) out[i] = order[i];
    }
    return (int64_t)order.size();
}
