// Ray / scene intersection on the device: primitive roots, point membership, CSG programs and the
// brute-force closest-hit loop of Universe::trace_closest (src/universe/mod.rs:85-147).
//
// The reference evaluates CSG with lazily pulled iterator objects (shape.rs:204-497).  Here each
// CSG program (post-order node list) is evaluated EAGERLY by a small stack machine: every node
// produces its hit list "up to the first None", which is exactly the prefix a parent can ever
// observe because a cached None is never stepped over (util.rs:395-419).  Hits are kept compact
// -- (t, primitive, flags) -- and locations / normals are recomputed from the primitive when
// needed, with the same expressions the reference uses, so results are bit-identical.
#pragma once
#include "scene_dev.cuh"
#include "vecmath.cuh"

namespace eucl {

constexpr int CSG_ARENA = 128;     // compact hits per thread (scene_create validates programs against it)
constexpr int CSG_LIST_STACK = 16; // nesting depth of pending child lists

struct CHit {
    double t;
    int prim;
    int flags; // bit 0: second root of the primitive; bit 1: normal negated (Complement / SymDiff)
};

// Roots of a*t^2 + b*t + c with the reference's selection of non-negative roots
// (shape.rs:672-694 sphere, :969-991 cylinder).  Returns the number of hits.
__device__ __forceinline__ int quadratic_hits(double a, double b, double c, double& t_first, double& t_second) {
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return 0;
    double d_sqrt = sqrt(disc);
    double t1 = (-b - d_sqrt) / (2.0 * a);
    double t2 = (-b + d_sqrt) / (2.0 * a);
    if (t1 >= 0.0) {
        t_first = t1;
        if (t2 >= 0.0) {
            t_second = t2;
            return 2;
        }
        return 1;
    }
    if (t2 >= 0.0) {
        t_first = t2;
        return 1;
    }
    return 0;
}

// Parametric distances at which the ray meets primitive `prim` (sorted, t >= 0 only).
template <int D>
__device__ __forceinline__ int prim_roots(const SceneView& sv, int prim, const Vec<D>& o, const Vec<D>& d, double& t0,
                                          double& t1) {
    const int n = sv.n_prims;
    const int kind = sv.prim_kind[prim];
    if (kind == EUCL_PRIM_SPHERE) { // shape.rs:662-670
        Vec<D> center = load_vec<D>(sv.prim_v0 + prim, n);
        double radius = sv.prim_s0[prim];
        Vec<D> rel = o - center;
        double a = norm_squared(d);
        double b = 2.0 * dot(d, rel);
        double c = norm_squared(rel) - radius * radius;
        return quadratic_hits(a, b, c, t0, t1);
    }
    if (kind == EUCL_PRIM_HYPERPLANE || kind == EUCL_PRIM_HALFSPACE) { // shape.rs:788-793
        Vec<D> nrm = load_vec<D>(sv.prim_v0 + prim, n);
        double t = -(dot(nrm, o) + sv.prim_s0[prim]) / dot(nrm, d);
        if (t < 0.0) return 0; // NaN and +inf pass, exactly like the reference
        t0 = t;
        return 1;
    }
    if (kind == EUCL_PRIM_CYLINDER) { // shape.rs:946-953
        Vec<D> center = load_vec<D>(sv.prim_v0 + prim, n);
        Vec<D> axis = load_vec<D>(sv.prim_v1 + prim, n);
        double radius = sv.prim_s0[prim];
        Vec<D> a_vec = d - axis * dot(d, axis);
        Vec<D> delta = o - center;
        Vec<D> c_vec = delta - axis * dot(delta, axis);
        double a = norm_squared(a_vec);
        double b = (1.0 + 1.0) * dot(a_vec, c_vec);
        double c = norm_squared(c_vec) - radius * radius;
        return quadratic_hits(a, b, c, t0, t1);
    }
    return 0; // VoidShape: shape.rs:622-631
}

// Shape::is_point_inside of one primitive (shape.rs:614-619,734-738,812-817,873-881,1030-1038)
template <int D>
__device__ __forceinline__ bool prim_inside(const SceneView& sv, int prim, const Vec<D>& p) {
    const int n = sv.n_prims;
    const int kind = sv.prim_kind[prim];
    if (kind == EUCL_PRIM_HALFSPACE) {
        double r = dot(load_vec<D>(sv.prim_v0 + prim, n), p) + sv.prim_s0[prim];
        return sv.prim_s1[prim] == rust_signum(r);
    }
    if (kind == EUCL_PRIM_SPHERE) {
        double radius = sv.prim_s0[prim];
        return norm_squared(load_vec<D>(sv.prim_v0 + prim, n) - p) <= radius * radius;
    }
    if (kind == EUCL_PRIM_CYLINDER) {
        Vec<D> center = load_vec<D>(sv.prim_v0 + prim, n);
        Vec<D> axis = load_vec<D>(sv.prim_v1 + prim, n);
        double radius = sv.prim_s0[prim];
        Vec<D> on_axis = axis * dot(axis, p - center) + center;
        return norm_squared(p - on_axis) <= radius * radius;
    }
    return kind == EUCL_PRIM_VOID; // hyperplane: never inside
}

// ComposableShape::is_point_inside (shape.rs:587-601) over the post-order range of node `n`,
// with a bit stack instead of recursion (all operands are pure, so no short-circuit is needed).
template <int D>
__device__ __forceinline__ bool node_inside(const SceneView& sv, int n, const Vec<D>& p) {
    const int first = sv.nodes[n].first;
    if (first == n) return prim_inside<D>(sv, sv.nodes[n].prim, p);
    unsigned long long bits = 0ull;
    int sp = 0;
    for (int m = first; m <= n; ++m) {
        const EuclNode nd = sv.nodes[m];
        if (nd.op == EUCL_CSG_LEAF) {
            bits |= (unsigned long long)(prim_inside<D>(sv, nd.prim, p) ? 1 : 0) << sp;
            ++sp;
        } else {
            bool b = (bits >> (sp - 1)) & 1ull, a = (bits >> (sp - 2)) & 1ull;
            bool r = nd.op == EUCL_CSG_UNION ? (a || b)
                     : nd.op == EUCL_CSG_INTERSECTION ? (a && b)
                     : nd.op == EUCL_CSG_COMPLEMENT ? (a && !b)
                                                    : (a != b);
            sp -= 2;
            bits &= ~(3ull << sp);
            bits |= (unsigned long long)(r ? 1 : 0) << sp;
            ++sp;
        }
    }
    return bits & 1ull;
}

// Universe::material_at (mod.rs:229-251): first entity in list order containing the point
template <int D>
__device__ __forceinline__ int material_at(const SceneView& sv, const Vec<D>& p) {
    for (int e = 0; e < sv.n_entities; ++e)
        if (node_inside<D>(sv, sv.entities[e].node_root, p)) return e;
    return -1;
}

// First item of the intersection stream of the CSG program [first, root]
// (ComposableShape::intersect_linear + the four merge iterators, shape.rs:204-584).
template <int D>
__device__ bool csg_first(const SceneView& sv, int first, int root, const Vec<D>& o, const Vec<D>& d, CHit& out) {
    CHit arena[CSG_ARENA];
    int lstart[CSG_LIST_STACK], llen[CSG_LIST_STACK];
    int sp = 0, top = 0;
    for (int n = first; n <= root; ++n) {
        const EuclNode nd = sv.nodes[n];
        if (nd.op == EUCL_CSG_LEAF) {
            double t0 = 0.0, t1 = 0.0;
            int c = prim_roots<D>(sv, nd.prim, o, d, t0, t1);
            lstart[sp] = top;
            llen[sp] = c;
            ++sp;
            if (c > 0) arena[top++] = CHit{t0, nd.prim, 0};
            if (c > 1) arena[top++] = CHit{t1, nd.prim, 1};
            continue;
        }
        const int b0 = lstart[sp - 1], bl = llen[sp - 1], a0 = lstart[sp - 2], al = llen[sp - 2];
        sp -= 2;
        const int nb = n - 1, na = sv.nodes[n - 1].first - 1;
        const int op = nd.op;
        const int out0 = top;
        // The entity root only ever yields its first item (mod.rs:110-112); inner nodes are
        // bounded because a Complement with an exhausted `b` repeats `a` forever (shape.rs:392).
        const int limit = n == root ? 1 : min(al + bl + 2, CSG_ARENA - out0);
        int outn = 0, ia = 0, ib = 0;
        while (outn < limit) {
            const bool has_a = ia < al, has_b = ib < bl;
            if (!has_a && !has_b) break;
            CHit ha = arena[a0 + (has_a ? ia : 0)], hb = arena[b0 + (has_b ? ib : 0)];
            if (has_a && has_b) {
                const bool a_closer = ha.t < hb.t; // ties (and NaN) pick b
                if (op == EUCL_CSG_COMPLEMENT) {
                    if (a_closer) {
                        ++ia;
                        if (!node_inside<D>(sv, nb, o + d * ha.t)) arena[out0 + outn++] = ha;
                    } else {
                        ++ib;
                        if (node_inside<D>(sv, na, o + d * hb.t)) {
                            hb.flags ^= 2;
                            arena[out0 + outn++] = hb;
                        }
                    }
                } else {
                    CHit closer = a_closer ? ha : hb;
                    const int further = a_closer ? nb : na;
                    if (a_closer) ++ia;
                    else ++ib;
                    const bool in = node_inside<D>(sv, further, o + d * closer.t);
                    if (op == EUCL_CSG_UNION) {
                        if (!in) arena[out0 + outn++] = closer;
                    } else if (op == EUCL_CSG_INTERSECTION) {
                        if (in) arena[out0 + outn++] = closer;
                    } else { // SymmetricDifference: always yields, flipped when inside the other
                        if (in) closer.flags ^= 2;
                        arena[out0 + outn++] = closer;
                    }
                }
            } else if (has_a) {
                if (op == EUCL_CSG_COMPLEMENT) {
                    arena[out0 + outn++] = ha; // not advanced (reference quirk)
                    continue;
                }
                ++ia;
                const bool in = node_inside<D>(sv, nb, o + d * ha.t);
                if (op == EUCL_CSG_UNION) {
                    if (in) break; // None
                } else if (op == EUCL_CSG_INTERSECTION) {
                    if (!in) break;
                } else if (in) {
                    ha.flags ^= 2;
                }
                arena[out0 + outn++] = ha;
            } else {
                ++ib;
                const bool in = node_inside<D>(sv, na, o + d * hb.t);
                if (op == EUCL_CSG_UNION) {
                    if (in) break;
                } else if (op == EUCL_CSG_INTERSECTION) {
                    if (!in) break;
                } else if (op == EUCL_CSG_COMPLEMENT) {
                    if (!in) break;
                    hb.flags ^= 2;
                } else if (in) {
                    hb.flags ^= 2;
                }
                arena[out0 + outn++] = hb;
            }
        }
        for (int k = 0; k < outn; ++k) arena[a0 + k] = arena[out0 + k];
        top = a0 + outn;
        lstart[sp] = a0;
        llen[sp] = outn;
        ++sp;
    }
    if (llen[0] > 0) {
        out = arena[lstart[0]];
        return true;
    }
    return false;
}

struct ClosestHit {
    int entity; // -1: no hit
    int prim;
    int flags;
    double t;
};

// trace_closest, distance part: every surfaced entity (including the one the ray is inside) is
// asked for the FIRST item of its stream; a candidate replaces the current one only if it is
// strictly closer (mod.rs:127-128), so a NaN distance wins only as the very first candidate.
template <int D>
__device__ __forceinline__ ClosestHit closest_hit(const SceneView& sv, const Vec<D>& o, const Vec<D>& d) {
    ClosestHit best{-1, 0, 0, 0.0};
    for (int e = 0; e < sv.n_entities; ++e) {
        const EuclEntity ent = sv.entities[e];
        if (ent.surface < 0) continue;
        CHit h;
        bool found;
        if (ent.node_first == ent.node_root) {
            double t0 = 0.0, t1 = 0.0;
            const int prim = sv.nodes[ent.node_root].prim;
            found = prim_roots<D>(sv, prim, o, d, t0, t1) > 0;
            h = CHit{t0, prim, 0};
        } else {
            found = csg_first<D>(sv, ent.node_first, ent.node_root, o, d, h);
        }
        if (found && (best.entity < 0 || best.t > h.t)) best = ClosestHit{e, h.prim, h.flags, h.t};
    }
    return best;
}

// Location and (raw) normal of a compact hit, as the primitive's intersector reports them
// (shape.rs:700-728 sphere, :795-806 plane, :858-861 half-space, :993-1024 cylinder).
template <int D>
__device__ __forceinline__ void hit_geometry(const SceneView& sv, int prim, int flags, const Vec<D>& o, const Vec<D>& d,
                                             double t, Vec<D>& p, Vec<D>& nrm) {
    const int n = sv.n_prims;
    const int kind = sv.prim_kind[prim];
    if (kind == EUCL_PRIM_SPHERE) {
        p = o + d * t;
        nrm = normalize(p - load_vec<D>(sv.prim_v0 + prim, n));
    } else if (kind == EUCL_PRIM_CYLINDER) {
        Vec<D> center = load_vec<D>(sv.prim_v0 + prim, n);
        Vec<D> axis = load_vec<D>(sv.prim_v1 + prim, n);
        double t_first = t;
        if (flags & 1) { // the second hit reuses the axis point of the FIRST hit (shape.rs:999,1017)
            double r0 = 0.0, r1 = 0.0;
            prim_roots<D>(sv, prim, o, d, r0, r1);
            t_first = r0;
        }
        Vec<D> p1 = o + d * t_first;
        Vec<D> on_axis = axis * dot(axis, p1 - center) + center;
        p = o + d * t;
        nrm = normalize(p - on_axis);
    } else {
        p = d * t + o;
        nrm = load_vec<D>(sv.prim_v0 + prim, n);
        if (kind == EUCL_PRIM_HALFSPACE) nrm = nrm * -sv.prim_s1[prim];
    }
    if (flags & 2) nrm = -nrm;
}

} // namespace eucl
