// Scene JSON -> flat tables.  Constructor names, argument order, named-argument keys, aliases and
// error classes follow the reference registry, src/scene.rs:618-1408; the object-or-array argument
// convention follows the `deserializer!` macro, src/scene.rs:457-515.
#include "scene_parse.h"

#include <cmath>
#include <cstring>
#include <functional>
#include <map>

#include "expr.h"
#include "json.h"

namespace eucl {

namespace {

// ---------------------------------------------------------------------------------------------
// intermediate objects (what the reference keeps as Box<dyn Trait>)

enum class Kind {
    Point, Vector, Rgba, Entity, Shape, SetOp, Material, CTExpr, LinTransform, UVFn, Texture,
    MappedTexture, Surface, BlendFn, ColorProvider, RatioProvider, ReflDirProvider,
    ThrDirProvider, Camera, Environment
};

const char* kind_name(Kind k) {
    switch (k) {
    case Kind::Point: return "Point";
    case Kind::Vector: return "Vector";
    case Kind::Rgba: return "Rgba";
    case Kind::Entity: return "Entity";
    case Kind::Shape: return "Shape";
    case Kind::SetOp: return "SetOperation";
    case Kind::Material: return "Material";
    case Kind::CTExpr: return "ComponentTransformationExpr";
    case Kind::LinTransform: return "LinearTransformation";
    case Kind::UVFn: return "UVFn";
    case Kind::Texture: return "Texture";
    case Kind::MappedTexture: return "MappedTexture";
    case Kind::Surface: return "Surface";
    case Kind::BlendFn: return "BlendFunction";
    case Kind::ColorProvider: return "SurfaceColorProvider";
    case Kind::RatioProvider: return "ReflectionRatioProvider";
    case Kind::ReflDirProvider: return "ReflectionDirectionProvider";
    case Kind::ThrDirProvider: return "ThresholdDirectionProvider";
    case Kind::Camera: return "Camera";
    case Kind::Environment: return "Environment";
    }
    return "?";
}

struct ShapeT {
    int op = EUCL_CSG_LEAF;
    EuclPrim prim{};
    std::shared_ptr<ShapeT> a, b;
};
using ShapeP = std::shared_ptr<ShapeT>;

struct ExprPair {
    std::string fwd, inv;
};
struct TransformT {
    std::vector<ExprPair> exprs;
};
struct MaterialT {
    int kind = EUCL_MAT_VACUUM;
    std::string legend;
    std::vector<TransformT> transforms;
};
struct TextureT {
    int filter = EUCL_TEX_LINEAR;
    std::string path;
};
struct MappedT {
    double center[4] = {0, 0, 0, 0};
    TextureT tex;
};
struct ColorT {
    int op = EUCL_COL_UNIFORM;
    double f[12] = {0};
    int blend = 0;
    std::shared_ptr<ColorT> src, dst;
    MappedT mapped;
};
using ColorP = std::shared_ptr<ColorT>;
struct SurfaceT {
    int ratio_op = 0;
    double ratio_a = 0, ratio_b = 0;
    int thr_op = 0;
    double thr_a = 0;
    ColorP color;
};
struct EntityT {
    ShapeP shape;
    MaterialT material;
    bool has_surface = false;
    SurfaceT surface;
};
struct EnvT {
    int dim = 3;
    EuclCamera camera{};
    std::vector<EntityT> entities;
    MappedT background;
};

struct Obj {
    Kind kind = Kind::Point;
    int dim = 0;        // 0 = dimension independent
    double v[4] = {0};  // Point / Vector / Rgba
    int i = 0;          // SetOp / BlendFn / ratio op / thr op
    double a = 0, b = 0; // provider parameters
    ShapeP shape;
    MaterialT material;
    ExprPair expr;
    TransformT transform;
    TextureT texture;
    MappedT mapped; // UVFn (center only) / MappedTexture
    ColorP color;
    SurfaceT surface;
    EntityT entity;
    EuclCamera camera{};
    std::shared_ptr<EnvT> env;
};

struct ParseError {
    int status;
    std::string message;
};

class Parser;
using Ctor = std::function<Obj(const JsonValue& parent, const JsonValue& json, Parser& parser)>;

// ---------------------------------------------------------------------------------------------
// vector helpers, written in the order nalgebra 0.8 evaluates them

double dotn(const double* a, const double* b, int d) {
    double s = a[0] * b[0];
    for (int k = 1; k < d; ++k) s = s + a[k] * b[k];
    return s;
}

void normalize_into(const double* v, int d, double* out) {
    double n = std::sqrt(dotn(v, v, d));
    for (int k = 0; k < d; ++k) out[k] = v[k] / n;
}

// ---------------------------------------------------------------------------------------------

class Parser {
public:
    Parser() { register_all(); }

    Obj deserialize_constructor(const JsonValue& json, Kind kind, int dim) {
        // src/scene.rs:1449-1464
        if (json.kind != JsonValue::Object || json.entries.size() != 1) {
            throw ParseError{EUCL_ERR_PARSE_INVALID_CONSTRUCTOR,
                             "A constructor must be an object containing a single key pointing to either an "
                             "object or an array. " + json.dump()};
        }
        const std::string& key = json.entries[0].first;
        const JsonValue& value = json.entries[0].second;
        auto it = ctors_.find(key);
        if (it == ctors_.end()) {
            throw ParseError{EUCL_ERR_PARSE_NO_DESERIALIZER, "No deserializer registered for key `" + key + "`."};
        }
        Obj result = it->second(json, value, *this);
        if (result.kind != kind || (dim != 0 && result.dim != 0 && result.dim != dim)) {
            throw ParseError{EUCL_ERR_PARSE_TYPE_MISMATCH,
                             "The constructor used (`" + key + "`) has an incorrect type for this field (expected " +
                                 kind_name(kind) + (dim ? std::to_string(dim) : std::string()) + "). " + json.dump()};
        }
        return result;
    }

private:
    std::map<std::string, Ctor> ctors_;

    void add(std::initializer_list<const char*> names, Ctor c) {
        for (const char* n : names) ctors_[n] = c;
    }

    // Argument reader: positional from an array, by key from an object (src/scene.rs:462-513).
    struct Args {
        const JsonValue& parent;
        const JsonValue& json;
        Parser& parser;
        size_t next = 0;

        Args(const JsonValue& p, const JsonValue& j, Parser& ps) : parent(p), json(j), parser(ps) {
            if (json.kind != JsonValue::Object && json.kind != JsonValue::Array) {
                throw ParseError{EUCL_ERR_PARSE_INVALID_CONSTRUCTOR,
                                 "The constructor data may only be an array or an object, received " + json.dump() +
                                     " instead."};
            }
        }
        const JsonValue& field(const char* name, const char* type_name) {
            if (json.kind == JsonValue::Object) {
                const JsonValue* v = json.get(name);
                if (!v) {
                    throw ParseError{EUCL_ERR_PARSE_MISSING_FIELD, std::string("Missing field of type ") + type_name +
                                                                       " with key " + name + " in " + parent.dump() + "."};
                }
                return *v;
            }
            if (next >= json.items.size()) {
                throw ParseError{EUCL_ERR_PARSE_MISSING_FIELD,
                                 std::string("Missing field of type ") + type_name + " in " + parent.dump() +
                                     ". To fix this, add the field at the end of the array."};
            }
            return json.items[next++];
        }
        [[noreturn]] void mismatch(const char* type_name, const JsonValue& v) {
            throw ParseError{EUCL_ERR_PARSE_TYPE_MISMATCH,
                             std::string("Expected `") + type_name + "`, could not parse from `" + v.dump() + "`."};
        }
        double f(const char* name) {
            const JsonValue& v = field(name, "F");
            if (v.kind != JsonValue::Number) mismatch("floating point number", v);
            return v.as_f64();
        }
        uint64_t u(const char* name, uint64_t max, const char* type_name) {
            const JsonValue& v = field(name, type_name);
            uint64_t r = 0;
            if (!v.as_u64(&r) || r > max) mismatch(type_name, v);
            return r;
        }
        std::string s(const char* name) {
            const JsonValue& v = field(name, "&str");
            if (v.kind != JsonValue::String) mismatch("string", v);
            return v.str;
        }
        Obj obj(const char* name, Kind kind, int dim) {
            return parser.deserialize_constructor(field(name, kind_name(kind)), kind, dim);
        }
        std::vector<Obj> vec(const char* name, Kind kind, int dim) {
            const JsonValue& v = field(name, "Vec");
            std::vector<Obj> out;
            // json crate: `.members()` of a non-array is an empty iterator
            if (v.kind == JsonValue::Array)
                for (const JsonValue& m : v.items) out.push_back(parser.deserialize_constructor(m, kind, dim));
            return out;
        }
    };

    // --- shape constructors (shape.rs) -------------------------------------------------------

    static ShapeP leaf(const EuclPrim& p) {
        auto s = std::make_shared<ShapeT>();
        s->op = EUCL_CSG_LEAF;
        s->prim = p;
        return s;
    }

    // Hyperplane::new (shape.rs:750-759)
    static EuclPrim hyperplane_new(const double* normal, double constant, int d) {
        if (!(dotn(normal, normal, d) > 0.0))
            throw ParseError{EUCL_ERR_PARSE_CUSTOM, "Cannot have a normal with length of 0."};
        EuclPrim p{};
        p.kind = EUCL_PRIM_HYPERPLANE;
        for (int k = 0; k < d; ++k) p.v0[k] = normal[k];
        p.s0 = constant;
        return p;
    }
    // Hyperplane::new_with_point (shape.rs:761-766)
    static EuclPrim hyperplane_with_point(const double* normal, const double* point, int d) {
        return hyperplane_new(normal, -dotn(normal, point, d), d);
    }
    // HalfSpace::new (shape.rs:828-835)
    static EuclPrim halfspace_new(const EuclPrim& plane, double signum) {
        EuclPrim p = plane;
        p.kind = EUCL_PRIM_HALFSPACE;
        p.s1 = signum / std::fabs(signum);
        return p;
    }
    // HalfSpace::new_with_point (shape.rs:837-841)
    static EuclPrim halfspace_with_point(const EuclPrim& plane, const double* inside, int d) {
        double identifier = dotn(plane.v0, inside, d) + plane.s0;
        return halfspace_new(plane, identifier);
    }
    // Cylinder::new (shape.rs:893-904)
    static EuclPrim cylinder_new(const double* center, const double* direction, double radius, int d) {
        if (!(dotn(direction, direction, d) > 0.0))
            throw ParseError{EUCL_ERR_PARSE_CUSTOM, "Cannot have a direction with length of 0."};
        if (!(radius > 0.0)) throw ParseError{EUCL_ERR_PARSE_CUSTOM, "The radius must be positive."};
        EuclPrim p{};
        p.kind = EUCL_PRIM_CYLINDER;
        for (int k = 0; k < d; ++k) p.v0[k] = center[k];
        normalize_into(direction, d, p.v1);
        p.s0 = radius;
        return p;
    }
    // ComposableShape::of -- left fold (shape.rs:523-545)
    static ShapeP compose(const std::vector<ShapeP>& shapes, int op) {
        if (shapes.size() < 2)
            throw ParseError{EUCL_ERR_PARSE_CUSTOM, "2 or more `Shape`s are needed to construct a `ComposableShape`."};
        ShapeP result;
        for (size_t k = 1; k < shapes.size(); ++k) {
            auto n = std::make_shared<ShapeT>();
            n->op = op;
            n->a = k == 1 ? shapes[0] : result;
            n->b = shapes[k];
            result = n;
        }
        return result;
    }
    // Cylinder::new_with_height (shape.rs:906-927)
    static ShapeP cylinder_with_height(const double* center, const double* direction, double radius, double height,
                                       int d) {
        double nd[4], top[4], bottom[4];
        normalize_into(direction, d, nd);
        double half_height = height / (1.0 + 1.0);
        for (int k = 0; k < d; ++k) {
            top[k] = center[k] + nd[k] * half_height;
            bottom[k] = center[k] + nd[k] * -half_height;
        }
        std::vector<ShapeP> shapes;
        shapes.push_back(leaf(cylinder_new(center, direction, radius, d)));
        shapes.push_back(leaf(halfspace_with_point(hyperplane_with_point(nd, top, d), center, d)));
        shapes.push_back(leaf(halfspace_with_point(hyperplane_with_point(nd, bottom, d), center, d)));
        return compose(shapes, EUCL_CSG_INTERSECTION);
    }
    static void cross3(const double* a, const double* b, double* out) {
        out[0] = a[1] * b[2] - a[2] * b[1];
        out[1] = a[2] * b[0] - a[0] * b[2];
        out[2] = a[0] * b[1] - a[1] * b[0];
    }
    // cuboid (d3/entity/shape.rs:17-66)
    static ShapeP cuboid(const double* center, const double* abc) {
        double half[3] = {abc[0] / 2.0, abc[1] / 2.0, abc[2] / 2.0};
        const double axes[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        // plane normals: cross(y,z), cross(x,z), cross(x,y)
        const int na[3] = {1, 0, 0}, nb[3] = {2, 2, 1};
        std::vector<ShapeP> shapes;
        for (int axis = 0; axis < 3; ++axis) {
            double normal[3];
            cross3(axes[na[axis]], axes[nb[axis]], normal);
            for (int sign = 0; sign < 2; ++sign) {
                double point[3];
                for (int k = 0; k < 3; ++k) {
                    double off = axes[axis][k] * half[k];
                    if (sign) off = -off;
                    point[k] = center[k] + off;
                }
                shapes.push_back(leaf(halfspace_with_point(hyperplane_with_point(normal, point, 3), center, 3)));
            }
        }
        return compose(shapes, EUCL_CSG_INTERSECTION);
    }
    // hypercuboid (d4/entity/shape.rs:18-76)
    static ShapeP hypercuboid(const double* center, const double* abcd) {
        double half[4];
        for (int k = 0; k < 4; ++k) half[k] = abcd[k] / 2.0;
        std::vector<ShapeP> shapes;
        for (int axis = 0; axis < 4; ++axis) {
            double e[4] = {0, 0, 0, 0}, normal[4];
            e[axis] = 1.0;
            normalize_into(e, 4, normal);
            for (int sign = 0; sign < 2; ++sign) {
                double point[4];
                for (int k = 0; k < 4; ++k) {
                    double off = e[k] * half[k];
                    if (sign) off = -off;
                    point[k] = center[k] + off;
                }
                shapes.push_back(leaf(halfspace_with_point(hyperplane_with_point(normal, point, 4), center, 4)));
            }
        }
        return compose(shapes, EUCL_CSG_INTERSECTION);
    }

    static EuclPrim expect_hyperplane(const Obj& o, int d) {
        if (!o.shape || o.shape->op != EUCL_CSG_LEAF || o.shape->prim.kind != EUCL_PRIM_HYPERPLANE)
            throw ParseError{EUCL_ERR_PARSE_CUSTOM,
                             std::string("Invalid type, expected a `Hyperplane") + std::to_string(d) + "`."};
        return o.shape->prim;
    }

    static Obj shape_obj(ShapeP s, int d) {
        Obj o;
        o.kind = Kind::Shape;
        o.dim = d;
        o.shape = std::move(s);
        return o;
    }

    static EuclCamera default_camera(int d) {
        // d3/entity/camera.rs:42-52, d4/entity/camera.rs:47-59
        EuclCamera c{};
        c.dim = d;
        c.max_depth = 10;
        c.fov_deg = 90;
        c.forward[0] = 1.0;
        c.up[2] = 1.0;
        c.left[1] = 1.0;
        return c;
    }

    void register_all() {
        // --- General (src/scene.rs:618-667) ---
        for (int d = 3; d <= 4; ++d) {
            std::string n = std::to_string(d);
            auto vec_ctor = [d](Kind kind) {
                return [d, kind](const JsonValue& p, const JsonValue& j, Parser& ps) {
                    Args a(p, j, ps);
                    Obj o;
                    o.kind = kind;
                    o.dim = d;
                    static const char* names[4] = {"x", "y", "z", "w"};
                    for (int k = 0; k < d; ++k) o.v[k] = a.f(names[k]);
                    return o;
                };
            };
            ctors_["Point" + n] = ctors_["Point" + n + "::new"] = vec_ctor(Kind::Point);
            ctors_["Vector" + n] = ctors_["Vector" + n + "::new"] = vec_ctor(Kind::Vector);
        }
        add({"Rgba", "Rgba::new"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj o;
            o.kind = Kind::Rgba;
            o.v[0] = a.f("r");
            o.v[1] = a.f("g");
            o.v[2] = a.f("b");
            o.v[3] = a.f("a");
            return o;
        });
        add({"Rgba::new_u8"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj o;
            o.kind = Kind::Rgba;
            static const char* names[4] = {"r", "g", "b", "a"};
            for (int k = 0; k < 4; ++k) o.v[k] = (double)a.u(names[k], 255, "u8") / 255.0; // palette new_u8
            return o;
        });
        add({"Rgba::from_hsva"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj o;
            o.kind = Kind::Rgba;
            double hue = a.f("hue"), sat = a.f("saturation"), val = a.f("value"), alpha = a.f("alpha");
            hsv_to_rgb(hue, sat, val, o.v);
            o.v[3] = alpha;
            return o;
        });

        for (int d = 3; d <= 4; ++d) {
            std::string n = std::to_string(d);
            // --- Entities (src/scene.rs:669-737) ---
            ctors_["Void" + n] = ctors_["Void" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Entity;
                o.dim = d;
                o.entity.material = a.obj("material", Kind::Material, d).material;
                EuclPrim v{};
                v.kind = EUCL_PRIM_VOID;
                o.entity.shape = leaf(v);
                return o;
            };
            ctors_["Void" + n + "::new_with_vacuum"] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Entity;
                o.dim = d;
                EuclPrim v{};
                v.kind = EUCL_PRIM_VOID;
                o.entity.shape = leaf(v);
                return o;
            };
            Ctor with_surface = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Entity;
                o.dim = d;
                o.entity.shape = a.obj("shape", Kind::Shape, d).shape;
                o.entity.material = a.obj("material", Kind::Material, d).material;
                o.entity.surface = a.obj("surface", Kind::Surface, d).surface;
                o.entity.has_surface = true;
                return o;
            };
            ctors_["Entity" + n + "Impl"] = ctors_["Entity" + n + "Impl::new"] =
                ctors_["Entity" + n + "Impl::new_with_surface"] = with_surface;
            ctors_["Entity" + n + "Impl::new_without_surface"] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Entity;
                o.dim = d;
                o.entity.shape = a.obj("shape", Kind::Shape, d).shape;
                o.entity.material = a.obj("material", Kind::Material, d).material;
                return o;
            };

            // --- Shapes (src/scene.rs:739-942) ---
            ctors_["VoidShape" + n] = ctors_["VoidShape" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                                Parser& ps) {
                Args a(p, j, ps);
                EuclPrim v{};
                v.kind = EUCL_PRIM_VOID;
                return shape_obj(leaf(v), d);
            };
            ctors_["ComposableShape" + n] = ctors_["ComposableShape" + n + "::new"] =
                ctors_["ComposableShape" + n + "::of"] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                    Args a(p, j, ps);
                    std::vector<Obj> shapes = a.vec("shapes", Kind::Shape, d);
                    int op = a.obj("operation", Kind::SetOp, 0).i;
                    std::vector<ShapeP> list;
                    for (auto& s : shapes) list.push_back(s.shape);
                    return shape_obj(compose(list, op), d);
                };
            ctors_["Sphere" + n] = ctors_["Sphere" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                          Parser& ps) {
                Args a(p, j, ps);
                Obj c = a.obj("center", Kind::Point, d);
                EuclPrim s{};
                s.kind = EUCL_PRIM_SPHERE;
                for (int k = 0; k < d; ++k) s.v0[k] = c.v[k];
                s.s0 = a.f("radius");
                return shape_obj(leaf(s), d);
            };
            ctors_["Hyperplane" + n] = ctors_["Hyperplane" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                                  Parser& ps) {
                Args a(p, j, ps);
                Obj normal = a.obj("normal", Kind::Vector, d);
                double constant = a.f("constant");
                return shape_obj(leaf(hyperplane_new(normal.v, constant, d)), d);
            };
            ctors_["Hyperplane" + n + "::new_with_point"] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj normal = a.obj("normal", Kind::Vector, d);
                Obj point = a.obj("point", Kind::Point, d);
                return shape_obj(leaf(hyperplane_with_point(normal.v, point.v, d)), d);
            };
            ctors_["HalfSpace" + n] = ctors_["HalfSpace" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                                Parser& ps) {
                Args a(p, j, ps);
                Obj plane = a.obj("plane", Kind::Shape, d);
                double sign = a.f("sign");
                return shape_obj(leaf(halfspace_new(expect_hyperplane(plane, d), sign)), d);
            };
            ctors_["HalfSpace" + n + "::new_with_point"] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj plane = a.obj("plane", Kind::Shape, d);
                Obj point = a.obj("point", Kind::Point, d);
                return shape_obj(leaf(halfspace_with_point(expect_hyperplane(plane, d), point.v, d)), d);
            };
            ctors_["Cylinder" + n] = ctors_["Cylinder" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                              Parser& ps) {
                Args a(p, j, ps);
                Obj center = a.obj("center", Kind::Point, d);
                Obj direction = a.obj("direction", Kind::Vector, d);
                double radius = a.f("radius");
                return shape_obj(leaf(cylinder_new(center.v, direction.v, radius, d)), d);
            };
            ctors_["Cylinder" + n + "::new_with_height"] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj center = a.obj("center", Kind::Point, d);
                Obj direction = a.obj("direction", Kind::Vector, d);
                double radius = a.f("radius");
                double height = a.f("height");
                return shape_obj(cylinder_with_height(center.v, direction.v, radius, height, d), d);
            };

            // --- Materials (src/scene.rs:944-1032) ---
            ctors_["Vacuum" + n] = ctors_["Vacuum" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                          Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Material;
                o.dim = d;
                return o;
            };
            ctors_["ComponentTransformation" + n] = ctors_["ComponentTransformation" + n + "::new"] =
                [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                    Args a(p, j, ps);
                    Obj o;
                    o.kind = Kind::LinTransform;
                    o.dim = d;
                    for (auto& e : a.vec("expressions", Kind::CTExpr, 0)) o.transform.exprs.push_back(e.expr);
                    return o;
                };
            ctors_["LinearSpace" + n] = ctors_["LinearSpace" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                                    Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Material;
                o.dim = d;
                o.material.kind = EUCL_MAT_LINEAR_SPACE;
                o.material.legend = a.s("legend");
                for (auto& t : a.vec("transformations", Kind::LinTransform, d))
                    o.material.transforms.push_back(t.transform);
                // material.rs:76-97: the reference panics on the first transition otherwise
                if ((int)o.material.legend.size() < d && !o.material.transforms.empty())
                    throw ParseError{EUCL_ERR_PARSE_CUSTOM, "The legend is too short! Make sure it is sufficient for " +
                                                                std::to_string(d) + " dimensions."};
                for (auto& t : o.material.transforms)
                    if ((int)t.exprs.size() != d)
                        throw ParseError{EUCL_ERR_PARSE_CUSTOM,
                                         "The number of functions must be equal to the number of dimensions (" +
                                             std::to_string(d) + ")!"};
                return o;
            };

            // --- Surfaces (src/scene.rs:1034-1334) ---
            ctors_["MappedTextureImpl" + n] = ctors_["MappedTextureImpl" + n + "::new"] =
                [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                    Args a(p, j, ps);
                    Obj o;
                    o.kind = Kind::MappedTexture;
                    o.dim = d;
                    o.mapped = a.obj("uvfn", Kind::UVFn, d).mapped;
                    o.mapped.tex = a.obj("texture", Kind::Texture, 0).texture;
                    return o;
                };
            ctors_["ComposableSurface" + n] = ctors_["ComposableSurface" + n + "::new"] =
                [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                    Args a(p, j, ps);
                    Obj o;
                    o.kind = Kind::Surface;
                    o.dim = d;
                    Obj ratio = a.obj("reflection_ratio", Kind::RatioProvider, d);
                    a.obj("reflection_direction", Kind::ReflDirProvider, d); // only specular exists
                    Obj thr = a.obj("threshold_direction", Kind::ThrDirProvider, d);
                    Obj color = a.obj("surface_color", Kind::ColorProvider, d);
                    o.surface.ratio_op = ratio.i;
                    o.surface.ratio_a = ratio.a;
                    o.surface.ratio_b = ratio.b;
                    o.surface.thr_op = thr.i;
                    o.surface.thr_a = thr.a;
                    o.surface.color = color.color;
                    return o;
                };
            auto color_obj = [d](ColorP c) {
                Obj o;
                o.kind = Kind::ColorProvider;
                o.dim = d;
                o.color = std::move(c);
                return o;
            };
            ctors_["surface_color_blend_" + n] = [d, color_obj](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                auto c = std::make_shared<ColorT>();
                c->op = EUCL_COL_BLEND;
                c->src = a.obj("source", Kind::ColorProvider, d).color;
                c->dst = a.obj("destination", Kind::ColorProvider, d).color;
                Obj fn = a.obj("blend_function", Kind::BlendFn, 0);
                c->blend = fn.i;
                c->f[0] = fn.a;
                return color_obj(c);
            };
            ctors_["surface_color_illumination_global_" + n] = [color_obj](const JsonValue& p, const JsonValue& j,
                                                                            Parser& ps) {
                Args a(p, j, ps);
                auto c = std::make_shared<ColorT>();
                c->op = EUCL_COL_ILLUM_GLOBAL;
                Obj light = a.obj("light_color", Kind::Rgba, 0);
                Obj dark = a.obj("dark_color", Kind::Rgba, 0);
                for (int k = 0; k < 4; ++k) {
                    c->f[k] = light.v[k];
                    c->f[4 + k] = dark.v[k];
                }
                return color_obj(c);
            };
            ctors_["surface_color_illumination_directional_" + n] = [d, color_obj](const JsonValue& p,
                                                                                    const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                auto c = std::make_shared<ColorT>();
                c->op = EUCL_COL_ILLUM_DIR;
                Obj dir = a.obj("direction", Kind::Vector, d);
                Obj light = a.obj("light_color", Kind::Rgba, 0);
                Obj dark = a.obj("dark_color", Kind::Rgba, 0);
                for (int k = 0; k < 4; ++k) {
                    c->f[k] = light.v[k];
                    c->f[4 + k] = dark.v[k];
                    c->f[8 + k] = k < d ? dir.v[k] : 0.0;
                }
                return color_obj(c);
            };
            ctors_["surface_color_uniform_" + n] = [color_obj](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                auto c = std::make_shared<ColorT>();
                c->op = EUCL_COL_UNIFORM;
                Obj color = a.obj("color", Kind::Rgba, 0);
                for (int k = 0; k < 4; ++k) c->f[k] = color.v[k];
                return color_obj(c);
            };
            ctors_["surface_color_texture_" + n] = [d, color_obj](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                auto c = std::make_shared<ColorT>();
                c->op = EUCL_COL_TEXTURE;
                c->mapped = a.obj("mapped_texture", Kind::MappedTexture, d).mapped;
                return color_obj(c);
            };
            ctors_["reflection_ratio_uniform_" + n] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::RatioProvider;
                o.dim = d;
                o.i = EUCL_RATIO_UNIFORM;
                o.a = a.f("ratio");
                return o;
            };
            ctors_["reflection_ratio_fresnel_" + n] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::RatioProvider;
                o.dim = d;
                o.i = EUCL_RATIO_FRESNEL;
                o.a = a.f("refractive_index_inside");
                o.b = a.f("refractive_index_outside");
                return o;
            };
            ctors_["reflection_direction_specular_" + n] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::ReflDirProvider;
                o.dim = d;
                o.i = EUCL_REFL_SPECULAR;
                return o;
            };
            ctors_["threshold_direction_snell_" + n] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::ThrDirProvider;
                o.dim = d;
                o.i = EUCL_THR_SNELL;
                o.a = a.f("refractive_index");
                return o;
            };
            ctors_["threshold_direction_identity_" + n] = [d](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::ThrDirProvider;
                o.dim = d;
                o.i = EUCL_THR_IDENTITY;
                return o;
            };

            // --- Environments (src/scene.rs:1336-1408) ---
            ctors_["Universe" + n] = ctors_["Universe" + n + "::new"] = [d](const JsonValue& p, const JsonValue& j,
                                                                              Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Environment;
                o.env = std::make_shared<EnvT>();
                o.env->dim = d;
                o.env->camera = a.obj("camera", Kind::Camera, d).camera;
                for (auto& e : a.vec("entities", Kind::Entity, d)) o.env->entities.push_back(e.entity);
                o.env->background = a.obj("background", Kind::MappedTexture, d).mapped;
                return o;
            };
        }

        // 3-D only constructors
        add({"Hyperplane3::new_with_vectors"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj first = a.obj("first", Kind::Vector, 3);
            Obj second = a.obj("second", Kind::Vector, 3);
            Obj point = a.obj("point", Kind::Point, 3);
            double normal[3];
            cross3(first.v, second.v, normal); // shape.rs:768-776
            return shape_obj(leaf(hyperplane_with_point(normal, point.v, 3)), 3);
        });
        add({"HalfSpace3::cuboid"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj center = a.obj("center", Kind::Point, 3);
            Obj dims = a.obj("dimensions", Kind::Vector, 3);
            return shape_obj(cuboid(center.v, dims.v), 3);
        });
        add({"HalfSpace4::hypercuboid"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj center = a.obj("center", Kind::Point, 4);
            Obj dims = a.obj("dimensions", Kind::Vector, 4);
            return shape_obj(hypercuboid(center.v, dims.v), 4);
        });
        add({"SetOperation", "SetOperation::new"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            std::string name = a.s("name");
            Obj o;
            o.kind = Kind::SetOp;
            if (name == "Union") o.i = EUCL_CSG_UNION;
            else if (name == "Intersection") o.i = EUCL_CSG_INTERSECTION;
            else if (name == "Complement") o.i = EUCL_CSG_COMPLEMENT;
            else if (name == "SymmetricDifference") o.i = EUCL_CSG_SYMDIFF;
            else throw ParseError{EUCL_ERR_PARSE_CUSTOM, "Invalid `SetOperation`: \"" + name + "\""};
            return o;
        });
        add({"ComponentTransformationExpr", "ComponentTransformationExpr::new"},
            [](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::CTExpr;
                o.expr.fwd = a.s("expression");
                o.expr.inv = a.s("inverse_expression");
                // syntax check now (Expr::from_str, src/scene.rs:964-981); variables bind later
                for (const std::string* e : {&o.expr.fwd, &o.expr.inv}) {
                    std::vector<EuclExprOp> scratch;
                    std::string err;
                    if (!expr_compile(*e, "abcdefghijklmnopqrstuvwxyz", 26, &scratch, &err))
                        throw ParseError{EUCL_ERR_PARSE_CUSTOM,
                                         "Invalid component transformation expression `" + *e + "`. (" + err + ")"};
                }
                return o;
            });
        add({"uv_sphere_3"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj o;
            o.kind = Kind::UVFn;
            o.dim = 3;
            Obj c = a.obj("center", Kind::Point, 3);
            for (int k = 0; k < 3; ++k) o.mapped.center[k] = c.v[k];
            return o;
        });
        add({"uv_derank_4"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj o = a.obj("uvfn", Kind::UVFn, 3);
            o.dim = 4; // d4/entity/surface.rs:11-15: drop w, then the 3-D mapping
            return o;
        });
        for (int filter = 0; filter < 2; ++filter) {
            const char* name = filter == EUCL_TEX_NEAREST ? "texture_image_nearest_neighbor" : "texture_image_linear";
            add({name}, [filter](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Texture;
                o.texture.filter = filter;
                o.texture.path = a.s("path");
                return o;
            });
        }
        add({"blend_function_ratio"}, [](const JsonValue& p, const JsonValue& j, Parser& ps) {
            Args a(p, j, ps);
            Obj o;
            o.kind = Kind::BlendFn;
            o.i = EUCL_BLEND_RATIO;
            o.a = a.f("ratio");
            return o;
        });
        static const struct {
            const char* name;
            int fn;
        } BLENDS[] = {{"over", EUCL_BLEND_OVER},           {"inside", EUCL_BLEND_INSIDE},
                      {"outside", EUCL_BLEND_OUTSIDE},     {"atop", EUCL_BLEND_ATOP},
                      {"xor", EUCL_BLEND_XOR},             {"plus", EUCL_BLEND_PLUS},
                      {"multiply", EUCL_BLEND_MULTIPLY},   {"screen", EUCL_BLEND_SCREEN},
                      {"overlay", EUCL_BLEND_OVERLAY},     {"darken", EUCL_BLEND_DARKEN},
                      {"lighten", EUCL_BLEND_LIGHTEN},     {"dodge", EUCL_BLEND_DODGE},
                      {"burn", EUCL_BLEND_BURN},           {"hard_light", EUCL_BLEND_HARD_LIGHT},
                      {"soft_light", EUCL_BLEND_SOFT_LIGHT}, {"difference", EUCL_BLEND_DIFFERENCE},
                      {"exclusion", EUCL_BLEND_EXCLUSION}};
        for (const auto& b : BLENDS) {
            int fn = b.fn;
            ctors_[std::string("blend_function_") + b.name] = [fn](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::BlendFn;
                o.i = fn;
                return o;
            };
        }
        // Perlin (3-D only, d3/entity/surface.rs:22-58).  The seed argument never reaches the
        // noise module in the reference (`perlin.set_seed(seed);` discards its result), so both
        // constructors lower to the default-seed table.
        auto perlin = [](bool seeded) {
            return [seeded](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                if (seeded) a.u("seed", 0xFFFFFFFFull, "u32");
                auto c = std::make_shared<ColorT>();
                c->op = EUCL_COL_PERLIN_HUE;
                c->f[0] = a.f("size");
                c->f[1] = a.f("speed");
                Obj o;
                o.kind = Kind::ColorProvider;
                o.dim = 3;
                o.color = c;
                return o;
            };
        };
        add({"surface_color_perlin_hue_seed_3"}, perlin(true));
        add({"surface_color_perlin_hue_random_3"}, perlin(false));

        // Cameras (src/scene.rs:1353-1408)
        auto camera = [](int d, bool with_location) {
            return [d, with_location](const JsonValue& p, const JsonValue& j, Parser& ps) {
                Args a(p, j, ps);
                Obj o;
                o.kind = Kind::Camera;
                o.dim = d;
                o.camera = default_camera(d);
                if (with_location) {
                    Obj loc = a.obj("location", Kind::Point, d);
                    for (int k = 0; k < d; ++k) o.camera.location[k] = loc.v[k];
                }
                return o;
            };
        };
        add({"PitchYawCamera3", "PitchYawCamera3::new", "FreeCamera3", "FreeCamera3::new"}, camera(3, false));
        add({"PitchYawCamera3::new_with_location", "FreeCamera3::new_with_location"}, camera(3, true));
        add({"FreeCamera4", "FreeCamera4::new"}, camera(4, false));
        add({"FreeCamera4::new_with_location"}, camera(4, true));
    }
};

// ---------------------------------------------------------------------------------------------
// lowering: EnvT -> flat tables

struct Lowerer {
    ParsedScene& out;
    int dim;

    int add_texture(const TextureT& t) {
        for (size_t k = 0; k < out.texture_paths.size(); ++k)
            if (out.texture_paths[k] == t.path) return (int)k;
        out.texture_paths.push_back(t.path);
        out.texture_pixels.emplace_back();
        out.texture_set.push_back(false);
        out.textures.push_back(EuclTexture{0, 0, 0});
        return (int)out.texture_paths.size() - 1;
    }
    int add_mapped(const MappedT& m) {
        EuclMappedTexture mt{};
        mt.uv_kind = EUCL_UV_SPHERE3;
        mt.filter = m.tex.filter;
        mt.texture = add_texture(m.tex);
        for (int k = 0; k < 4; ++k) mt.center[k] = m.center[k];
        out.mapped_textures.push_back(mt);
        return (int)out.mapped_textures.size() - 1;
    }
    void emit_shape(const ShapeP& s) {
        if (s->op == EUCL_CSG_LEAF) {
            out.prims.push_back(s->prim);
            int idx = (int)out.nodes.size();
            out.nodes.push_back(EuclNode{EUCL_CSG_LEAF, (int)out.prims.size() - 1, idx, 0});
            return;
        }
        int first = (int)out.nodes.size();
        emit_shape(s->a);
        emit_shape(s->b);
        out.nodes.push_back(EuclNode{s->op, -1, first, 0});
    }
    void emit_color(const ColorP& c) { // postfix: source, destination, BLEND
        EuclColorOp op{};
        op.op = c->op;
        for (int k = 0; k < 12; ++k) op.f[k] = c->f[k];
        if (c->op == EUCL_COL_BLEND) {
            emit_color(c->src);
            emit_color(c->dst);
            op.i0 = c->blend;
        } else if (c->op == EUCL_COL_TEXTURE) {
            op.i0 = add_mapped(c->mapped);
        }
        out.color_ops.push_back(op);
    }
    int add_material(const MaterialT& m) {
        EuclMaterial em{};
        em.kind = m.kind;
        em.transform_first = (int)out.transforms.size();
        em.n_transforms = (int)m.transforms.size();
        for (const TransformT& t : m.transforms) {
            EuclTransform et{};
            for (int k = 0; k < dim; ++k) {
                for (int inverse = 0; inverse < 2; ++inverse) {
                    const std::string& text = inverse ? t.exprs[k].inv : t.exprs[k].fwd;
                    int first = (int)out.expr_ops.size();
                    std::string err;
                    if (!expr_compile(text, m.legend, dim, &out.expr_ops, &err))
                        throw ParseError{EUCL_ERR_PARSE_CUSTOM,
                                         "Invalid component transformation expression `" + text + "`. (" + err + ")"};
                    int len = (int)out.expr_ops.size() - first;
                    (inverse ? et.inv_first : et.fwd_first)[k] = first;
                    (inverse ? et.inv_len : et.fwd_len)[k] = len;
                }
            }
            out.transforms.push_back(et);
        }
        out.materials.push_back(em);
        return (int)out.materials.size() - 1;
    }
    int add_surface(const SurfaceT& s) {
        EuclSurface es{};
        es.ratio_op = s.ratio_op;
        es.ratio_a = s.ratio_a;
        es.ratio_b = s.ratio_b;
        es.refl_op = EUCL_REFL_SPECULAR;
        es.thr_op = s.thr_op;
        es.thr_a = s.thr_a;
        es.color_first = (int)out.color_ops.size();
        emit_color(s.color);
        es.color_len = (int)out.color_ops.size() - es.color_first;
        out.surfaces.push_back(es);
        return (int)out.surfaces.size() - 1;
    }
    void run(const EnvT& env) {
        dim = env.dim;
        for (const EntityT& e : env.entities) {
            EuclEntity ee{};
            ee.node_first = (int)out.nodes.size();
            emit_shape(e.shape);
            ee.node_root = (int)out.nodes.size() - 1;
            ee.material = add_material(e.material);
            ee.surface = e.has_surface ? add_surface(e.surface) : -1;
            out.entities.push_back(ee);
        }
        out.flat.dim = dim;
        out.flat.background = add_mapped(env.background);
        out.flat.camera = env.camera;
        perlin_permutation(0, out.flat.perlin_perm);
        out.refresh_flat();
    }
};

} // namespace

void ParsedScene::refresh_flat() {
    texels.clear();
    for (size_t k = 0; k < textures.size(); ++k) {
        textures[k].texel_offset = texels.size();
        texels.insert(texels.end(), texture_pixels[k].begin(), texture_pixels[k].end());
    }
    flat.n_prims = (int)prims.size();
    flat.n_nodes = (int)nodes.size();
    flat.n_entities = (int)entities.size();
    flat.n_materials = (int)materials.size();
    flat.n_transforms = (int)transforms.size();
    flat.n_expr_ops = (int)expr_ops.size();
    flat.n_surfaces = (int)surfaces.size();
    flat.n_color_ops = (int)color_ops.size();
    flat.n_mapped_textures = (int)mapped_textures.size();
    flat.n_textures = (int)textures.size();
    flat.prims = prims.data();
    flat.nodes = nodes.data();
    flat.entities = entities.data();
    flat.materials = materials.data();
    flat.transforms = transforms.data();
    flat.expr_ops = expr_ops.data();
    flat.surfaces = surfaces.data();
    flat.color_ops = color_ops.data();
    flat.mapped_textures = mapped_textures.data();
    flat.textures = textures.data();
    flat.texels = texels.data();
    flat.texel_bytes = texels.size();
}

int parse_scene(const std::string& json_text, std::unique_ptr<ParsedScene>* out, std::string* error) {
    JsonValue root;
    std::string jerr;
    if (!json_parse(json_text, &root, &jerr)) {
        *error = "Invalid JSON file. Please, check the syntax. (" + jerr + ")";
        return EUCL_ERR_PARSE_SYNTAX;
    }
    try {
        Parser parser;
        Obj env = parser.deserialize_constructor(root, Kind::Environment, 0);
        auto scene = std::make_unique<ParsedScene>();
        Lowerer lower{*scene, env.env->dim};
        lower.run(*env.env);
        *out = std::move(scene);
        return EUCL_OK;
    } catch (const ParseError& e) {
        *error = e.message;
        return e.status;
    }
}

void perlin_permutation(uint32_t seed, uint8_t out[256]) {
    // rand 0.3/0.4 XorShiftRng::from_seed([1, seed, seed, seed]); seq = 0..=255; rng.shuffle(seq)
    uint32_t x = 1, y = seed, z = seed, w = seed;
    auto next_u32 = [&]() {
        uint32_t t = x ^ (x << 11);
        x = y;
        y = z;
        z = w;
        w = w ^ (w >> 19) ^ (t ^ (t >> 8));
        return w;
    };
    auto next_u64 = [&]() {
        uint64_t hi = next_u32();
        uint64_t lo = next_u32();
        return (hi << 32) | lo;
    };
    for (int k = 0; k < 256; ++k) out[k] = (uint8_t)k;
    size_t i = 256;
    while (i >= 2) {
        i -= 1;
        // gen_range(0, i + 1): 64-bit rejection sampling
        uint64_t range = (uint64_t)i + 1;
        uint64_t zone = UINT64_MAX - UINT64_MAX % range;
        uint64_t v;
        do {
            v = next_u64();
        } while (!(v < zone));
        size_t jdx = (size_t)(v % range);
        uint8_t tmp = out[i];
        out[i] = out[jdx];
        out[jdx] = tmp;
    }
}

void hsv_to_rgb(double hue_degrees, double saturation, double value, double rgb[3]) {
    // palette 0.2.1: RgbHue::to_positive_degrees, then `impl From<Hsv> for Rgb`
    double deg = hue_degrees;
    if (std::isfinite(deg)) {
        while (deg >= 360.0) deg = deg - 360.0;
        while (deg < 0.0) deg = deg + 360.0;
    }
    double c = value * saturation;
    double h = deg / 60.0;
    double x = c * (1.0 - std::fabs(std::fmod(h, 2.0) - 1.0));
    double m = value - c;
    double r, g, b;
    if (h >= 0.0 && h < 1.0) { r = c; g = x; b = 0.0; }
    else if (h >= 1.0 && h < 2.0) { r = x; g = c; b = 0.0; }
    else if (h >= 2.0 && h < 3.0) { r = 0.0; g = c; b = x; }
    else if (h >= 3.0 && h < 4.0) { r = 0.0; g = x; b = c; }
    else if (h >= 4.0 && h < 5.0) { r = x; g = 0.0; b = c; }
    else { r = c; g = 0.0; b = x; }
    rgb[0] = r + m;
    rgb[1] = g + m;
    rgb[2] = b + m;
}

} // namespace eucl
