// oracle/ikd_ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle, "reference" kind).
//
// Thin extern "C" wrapper around the UNMODIFIED reference ikd-Tree, compiled from
// the sources where they lie (/root/reference/eskf_lio/include/ikd-Tree/ikd_Tree.cpp)
// by oracle/Makefile into oracle/_ref/libikd_ref.so.  Nothing of the reference is
// copied into this repository; this file only calls its public methods:
//   Build               ikd_Tree.cpp:408     Nearest_Search   ikd_Tree.cpp:425
//   Add_Points          ikd_Tree.cpp:477     Delete_Point_Boxes ikd_Tree.cpp:631
//   flatten             ikd_Tree.cpp:1626    validnum / size  ikd_Tree.cpp:68,145
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the resulting library.
#include <ikd-Tree/ikd_Tree.h>

#include <cstring>
#include <vector>

typedef pcl::PointXYZINormal RefPoint;
typedef KD_TREE<RefPoint> RefTree;
typedef RefTree::PointVector RefVec;

extern "C" {

void *ikdref_create(float downsample_size) {
    RefTree *t = new RefTree();  // heap: the object embeds a 1e6-entry queue (ikd_Tree.h:18,207)
    t->set_downsample_param(downsample_size);
    return t;
}

void ikdref_destroy(void *h) { delete static_cast<RefTree *>(h); }

static inline RefPoint mk(const float *p) {
    RefPoint q;
    q.x = p[0];
    q.y = p[1];
    q.z = p[2];
    q.intensity = p[3];
    return q;
}

// pts: n x 4 floats (x y z intensity)
void ikdref_build(void *h, const float *pts, int n) {
    RefVec v(n);
    for (int i = 0; i < n; i++) v[i] = mk(pts + 4 * i);
    static_cast<RefTree *>(h)->Build(v);
}

// q: nq x 3 floats.  out_pts: nq x k x 4 floats, out_d: nq x k, out_cnt: nq.
void ikdref_knn(void *h, const float *q, int nq, int k, float *out_pts, float *out_d, int *out_cnt) {
    RefTree *t = static_cast<RefTree *>(h);
    RefVec near;
    std::vector<float> dist;
    for (int i = 0; i < nq; i++) {
        RefPoint p;
        p.x = q[3 * i];
        p.y = q[3 * i + 1];
        p.z = q[3 * i + 2];
        t->Nearest_Search(p, k, near, dist);
        int c = (int)near.size();
        out_cnt[i] = c;
        for (int j = 0; j < k; j++) {
            float *o = out_pts + ((size_t)i * k + j) * 4;
            if (j < c) {
                o[0] = near[j].x;
                o[1] = near[j].y;
                o[2] = near[j].z;
                o[3] = near[j].intensity;
                out_d[(size_t)i * k + j] = dist[j];
            } else {
                o[0] = o[1] = o[2] = o[3] = 0.f;
                out_d[(size_t)i * k + j] = -1.f;
            }
        }
    }
}

int ikdref_add(void *h, const float *pts, int n, int downsample_on) {
    if (n <= 0) return 0;
    RefVec v(n);
    for (int i = 0; i < n; i++) v[i] = mk(pts + 4 * i);
    return static_cast<RefTree *>(h)->Add_Points(v, downsample_on != 0);
}

// boxes: nb x 6 floats (min xyz, max xyz)
int ikdref_delete_boxes(void *h, const float *boxes, int nb) {
    std::vector<BoxPointType> b(nb);
    for (int i = 0; i < nb; i++) {
        for (int a = 0; a < 3; a++) {
            b[i].vertex_min[a] = boxes[6 * i + a];
            b[i].vertex_max[a] = boxes[6 * i + 3 + a];
        }
    }
    return static_cast<RefTree *>(h)->Delete_Point_Boxes(b);
}

int ikdref_validnum(void *h) { return static_cast<RefTree *>(h)->validnum(); }
int ikdref_size(void *h) { return static_cast<RefTree *>(h)->size(); }
int ikdref_has_root(void *h) { return static_cast<RefTree *>(h)->Root_Node != nullptr; }

// out: cap x 4 floats; returns number of live points (may exceed cap; only cap written)
int ikdref_flatten(void *h, float *out, int cap) {
    RefTree *t = static_cast<RefTree *>(h);
    RefVec v;
    if (t->Root_Node) t->flatten(t->Root_Node, v, NOT_RECORD);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; i++) {
        out[4 * i] = v[i].x;
        out[4 * i + 1] = v[i].y;
        out[4 * i + 2] = v[i].z;
        out[4 * i + 3] = v[i].intensity;
    }
    return n;
}

}  // extern "C"
