// rr_yaml.cpp — Scene YAML (de)serialisation, render.rs:668-676, 735-799 (schema: SURVEY.md appendix B).
//
// serde_yaml 0.8.11 is not vendored in the reference and there is no yaml library in this image, so
// this is a small hand-written reader for the YAML subset serde_yaml emits and people hand-edit
// (block mappings/sequences by indentation, flow [] / {} collections, plain / quoted scalars, `~`,
// comments, `---`), and a writer that follows serde_yaml 0.8's block style: `---` header, two-space
// indents, externally tagged enums (`- Sphere:`), f32 widened to f64 and printed shortest-round-trip.
// Format parity with the Rust binary is unpinned (no sample file exists); files written here load in
// PyYAML and in this reader (tests/test_host_cpu.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "rr_host.hpp"

namespace rr {
namespace {

struct Node {
    enum Type { Null, Scalar, Seq, Map } type = Null;
    std::string scalar;
    std::vector<Node> seq;
    std::vector<std::pair<std::string, Node>> map;  // insertion order
    const Node *get(const std::string &k) const {
        for (auto &kv : map) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Line {
    int indent;
    std::string text;  // without indent and trailing comment
};

std::string rstrip(const std::string &s) {
    size_t e = s.find_last_not_of(" \t\r\n");
    return e == std::string::npos ? "" : s.substr(0, e + 1);
}
std::string lstrip(const std::string &s) {
    size_t b = s.find_first_not_of(" \t");
    return b == std::string::npos ? "" : s.substr(b);
}

// remove a trailing `# comment` that is outside quotes
std::string strip_comment(const std::string &s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        if (q) {
            if (c == q) q = 0;
        } else if (c == '"' || c == '\'') {
            q = c;
        } else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) {
            return s.substr(0, i);
        }
    }
    return s;
}

std::vector<Line> split_lines(const std::string &text) {
    std::vector<Line> out;
    std::istringstream in(text);
    std::string raw;
    while (std::getline(in, raw)) {
        std::string s = rstrip(strip_comment(raw));
        if (s.empty()) continue;
        if (s == "---" || s == "...") continue;
        if (s.rfind("%", 0) == 0) continue;  // directives
        int ind = 0;
        while (ind < (int)s.size() && s[ind] == ' ') ++ind;
        if (ind < (int)s.size() && s[ind] == '\t') throw ParseError("tab indentation");
        out.push_back(Line{ind, s.substr(ind)});
    }
    return out;
}

std::string unquote(const std::string &s) {
    if (s.size() >= 2 && s.front() == '"' && s.back() == '"') {
        std::string o;
        for (size_t i = 1; i + 1 < s.size(); ++i) {
            if (s[i] == '\\' && i + 2 < s.size()) {
                char c = s[++i];
                o += c == 'n' ? '\n' : c == 't' ? '\t' : c;
            } else {
                o += s[i];
            }
        }
        return o;
    }
    if (s.size() >= 2 && s.front() == '\'' && s.back() == '\'') {
        std::string o;
        for (size_t i = 1; i + 1 < s.size(); ++i) {
            if (s[i] == '\'' && s[i + 1] == '\'' && i + 2 < s.size()) ++i;
            o += s[i];
        }
        return o;
    }
    return s;
}

// ---- flow collections: [a, b], {k: v, ...} -----------------------------------------------------
struct Flow {
    const std::string &s;
    size_t i = 0;
    explicit Flow(const std::string &s_) : s(s_) {}
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i; }
    Node value() {
        ws();
        if (i >= s.size()) return Node();
        if (s[i] == '[') {
            Node n; n.type = Node::Seq; ++i; ws();
            if (i < s.size() && s[i] == ']') { ++i; return n; }
            for (;;) {
                n.seq.push_back(value()); ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; break; }
                throw ParseError("bad flow sequence");
            }
            return n;
        }
        if (s[i] == '{') {
            Node n; n.type = Node::Map; ++i; ws();
            if (i < s.size() && s[i] == '}') { ++i; return n; }
            for (;;) {
                ws();
                std::string k = token(":");
                ws();
                if (i >= s.size() || s[i] != ':') throw ParseError("bad flow mapping");
                ++i;
                n.map.emplace_back(unquote(rstrip(k)), value()); ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; break; }
                throw ParseError("bad flow mapping");
            }
            return n;
        }
        std::string t = rstrip(token(",]}"));
        Node n;
        if (t == "~" || t == "null" || t.empty()) return n;
        n.type = Node::Scalar;
        n.scalar = unquote(t);
        return n;
    }
    std::string token(const char *stops) {
        size_t b = i;
        if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
            char q = s[i++];
            while (i < s.size() && s[i] != q) ++i;
            if (i < s.size()) ++i;
            return s.substr(b, i - b);
        }
        while (i < s.size() && !strchr(stops, s[i])) ++i;
        return s.substr(b, i - b);
    }
};

Node scalar_or_flow(const std::string &t) {
    std::string s = lstrip(t);
    if (s.empty() || s == "~" || s == "null") return Node();
    if (s[0] == '[' || s[0] == '{') {
        Flow f(s);
        return f.value();
    }
    Node n;
    n.type = Node::Scalar;
    n.scalar = unquote(s);
    return n;
}

// split "key: value" at the first ": " / trailing ":" outside quotes; returns false if not a mapping line
bool split_key(const std::string &t, std::string &key, std::string &rest) {
    char q = 0;
    for (size_t i = 0; i < t.size(); ++i) {
        char c = t[i];
        if (q) { if (c == q) q = 0; continue; }
        if (c == '"' || c == '\'') { q = c; continue; }
        if (c == '[' || c == '{') return false;
        if (c == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) {
            key = unquote(rstrip(t.substr(0, i)));
            rest = i + 1 < t.size() ? lstrip(t.substr(i + 1)) : "";
            return true;
        }
    }
    return false;
}

struct Parser {
    std::vector<Line> L;
    size_t pos = 0;

    Node block(int indent) {
        if (pos >= L.size() || L[pos].indent < indent) return Node();
        const int ind = L[pos].indent;
        if (L[pos].text[0] == '-' && (L[pos].text.size() == 1 || L[pos].text[1] == ' ')) return sequence(ind);
        std::string k, r;
        if (split_key(L[pos].text, k, r)) return mapping(ind);
        Node n = scalar_or_flow(L[pos].text);
        ++pos;
        return n;
    }
    Node sequence(int ind) {
        Node n; n.type = Node::Seq;
        while (pos < L.size() && L[pos].indent == ind && L[pos].text[0] == '-' &&
               (L[pos].text.size() == 1 || L[pos].text[1] == ' ')) {
            std::string rest = L[pos].text.size() > 1 ? L[pos].text.substr(2) : "";
            size_t extra = 0;
            while (extra < rest.size() && rest[extra] == ' ') ++extra;
            rest = rest.substr(extra);
            if (rest.empty()) {
                ++pos;
                n.seq.push_back(block(ind + 1));
            } else {
                // "- key: value" starts a nested block whose indent is the column of `key`
                std::string k, r;
                if (split_key(rest, k, r)) {
                    L[pos].indent = ind + 2 + (int)extra;
                    L[pos].text = rest;
                    n.seq.push_back(mapping(L[pos].indent));
                } else {
                    n.seq.push_back(scalar_or_flow(rest));
                    ++pos;
                }
            }
        }
        return n;
    }
    Node mapping(int ind) {
        Node n; n.type = Node::Map;
        while (pos < L.size() && L[pos].indent == ind) {
            std::string k, r;
            if (!split_key(L[pos].text, k, r)) throw ParseError("expected `key: value`");
            ++pos;
            if (!r.empty()) {
                n.map.emplace_back(k, scalar_or_flow(r));
            } else if (pos < L.size() && (L[pos].indent > ind || (L[pos].indent == ind && L[pos].text[0] == '-' &&
                                                                   (L[pos].text.size() == 1 || L[pos].text[1] == ' ')))) {
                n.map.emplace_back(k, block(L[pos].indent));  // nested block (sequences may sit at the key's indent)
            } else {
                n.map.emplace_back(k, Node());
            }
        }
        if (pos < L.size() && L[pos].indent > ind) throw ParseError("bad indentation");
        return n;
    }
};

Node parse_yaml(const std::string &text) {
    Parser p;
    p.L = split_lines(text);
    if (p.L.empty()) return Node();
    Node n = p.block(p.L[0].indent);
    if (p.pos != p.L.size()) throw ParseError("trailing content");
    return n;
}

// ---- typed accessors (serde: every field is required, render.rs has no #[serde(default)]) --------
const Node &field(const Node &m, const char *k) {
    if (m.type != Node::Map) throw ParseError(std::string("expected a mapping with field ") + k);
    const Node *n = m.get(k);
    if (!n) throw ParseError(std::string("missing field ") + k);
    return *n;
}
double as_f64(const Node &n) {
    if (n.type != Node::Scalar) throw ParseError("expected a number");
    const std::string &s = n.scalar;
    if (s == ".inf" || s == ".Inf" || s == ".INF" || s == "+.inf") return INFINITY;
    if (s == "-.inf" || s == "-.Inf" || s == "-.INF") return -INFINITY;
    if (s == ".nan" || s == ".NaN" || s == ".NAN") return NAN;
    char *end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != 0) throw ParseError("invalid number: " + s);
    return v;
}
float as_f32(const Node &n) { return (float)as_f64(n); }
int as_i32(const Node &n) {
    if (n.type != Node::Scalar) throw ParseError("expected an integer");
    char *end = nullptr;
    long v = strtol(n.scalar.c_str(), &end, 10);
    if (end == n.scalar.c_str() || *end != 0) throw ParseError("invalid integer: " + n.scalar);
    return (int)v;
}
std::string as_str(const Node &n) {
    if (n.type == Node::Null) return "";
    if (n.type != Node::Scalar) throw ParseError("expected a string");
    return n.scalar;
}
Vec3 as_vec3(const Node &n) { return Vec3(as_f32(field(n, "x")), as_f32(field(n, "y")), as_f32(field(n, "z"))); }
RenderColor as_color(const Node &n) { return RenderColor(as_f32(field(n, "r")), as_f32(field(n, "g")), as_f32(field(n, "b"))); }

template <typename E>
E as_enum(const Node &n, const std::vector<std::pair<const char *, E>> &names) {
    std::string s = as_str(n);
    for (auto &kv : names) if (s == kv.first) return kv.second;
    throw ParseError("unknown variant " + s);
}
const std::vector<std::pair<const char *, RenderPattern>> PATTERNS = {
    {"Solid", RenderPattern::Solid}, {"Checkerboard", RenderPattern::Checkerboard}, {"RepeatedGradation", RenderPattern::RepeatedGradation}};
const std::vector<std::pair<const char *, UVMap>> UVMAPS = {{"XY", UVMap::XY}, {"YZ", UVMap::YZ}, {"ZX", UVMap::ZX}, {"LL", UVMap::LL}};
const std::vector<std::pair<const char *, TextureFilter>> FILTERS = {{"Nearest", TextureFilter::Nearest}, {"Bilinear", TextureFilter::Bilinear}};
template <typename E>
const char *enum_name(E v, const std::vector<std::pair<const char *, E>> &names) {
    for (auto &kv : names) if (kv.second == v) return kv.first;
    return "?";
}

// ---- writer ---------------------------------------------------------------------------------------
// serde_yaml 0.8 widens f32 to f64 and prints the shortest string that round-trips the f64.
std::string fmt_f32(float v) {
    double d = (double)v;
    if (std::isnan(d)) return ".nan";
    if (std::isinf(d)) return d > 0 ? ".inf" : "-.inf";
    char buf[64];
    int prec = 1;
    for (; prec <= 17; ++prec) {
        snprintf(buf, sizeof buf, "%.*e", prec - 1, d);
        if (strtod(buf, nullptr) == d) break;
    }
    // ryu/dtoa style: positional notation for 1e-5 <= |d| < 1e16, exponent otherwise
    const double a = std::fabs(d);
    if (a == 0.0 || (a >= 1e-5 && a < 1e16)) {
        const int exp10 = a == 0.0 ? 0 : (int)std::floor(std::log10(a));
        int decimals = prec - 1 - exp10;
        if (decimals < 0) decimals = 0;
        snprintf(buf, sizeof buf, "%.*f", decimals, d);
        if (strtod(buf, nullptr) != d) snprintf(buf, sizeof buf, "%.*f", decimals + 1, d);  // log10 edge
    } else {
        snprintf(buf, sizeof buf, "%.*e", prec - 1, d);
    }
    std::string s = buf;
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}
std::string fmt_str(const std::string &s) {
    if (s.empty()) return "\"\"";
    bool plain = true;
    for (char c : s)
        if (!(isalnum((unsigned char)c) || c == '_' || c == '-' || c == '.' || c == '/')) plain = false;
    // strings that would re-parse as another type must be quoted
    char *end = nullptr;
    strtod(s.c_str(), &end);
    if (*end == 0 || s == "~" || s == "null" || s == "true" || s == "false") plain = false;
    if (plain) return s;
    std::string o = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\';
        o += c;
    }
    return o + "\"";
}
void emit_vec3(std::ostringstream &o, const std::string &pad, const char *key, const Vec3 &v) {
    o << pad << key << ":\n" << pad << "  x: " << fmt_f32(v.x) << "\n" << pad << "  y: " << fmt_f32(v.y) << "\n" << pad << "  z: " << fmt_f32(v.z) << "\n";
}
void emit_color(std::ostringstream &o, const std::string &pad, const char *key, const RenderColor &c) {
    o << pad << key << ":\n" << pad << "  r: " << fmt_f32(c.r) << "\n" << pad << "  g: " << fmt_f32(c.g) << "\n" << pad << "  b: " << fmt_f32(c.b) << "\n";
}

}  // namespace

// RenderEnv::serialize, render.rs:735-760
std::string RenderEnv::serialize() const {
    std::ostringstream o;
    o << "---\n";
    o << "camera:\n";
    emit_vec3(o, "  ", "position", camera.position);
    emit_vec3(o, "  ", "pyr", camera.pyr);
    o << "camera_motion: []\n";                         // always written empty (render.rs:741)
    o << "max_reflections: " << MAX_REFLECTIONS << "\n";  // the constants, not the env's values (render.rs:742-743)
    o << "max_refractions: " << MAX_REFRACTIONS << "\n";
    // materials of the OBJECTS keyed by name (render.rs:751-756); HashMap order is unspecified, sorted here
    std::map<std::string, const RenderMaterial *> mats;
    for (const RenderObject &ob : objects_) mats[ob.material->name_] = ob.material.get();
    if (mats.empty()) o << "materials: {}\n";
    else o << "materials:\n";
    for (auto &kv : mats) {
        const RenderMaterial &m = *kv.second;
        o << "  " << fmt_str(kv.first) << ":\n";
        o << "    name: " << fmt_str(m.name_) << "\n";
        emit_color(o, "    ", "diffuse", m.diffuse_);
        emit_color(o, "    ", "specular", m.specular_);
        o << "    pn: " << m.pn_ << "\n";
        o << "    t: " << fmt_f32(m.t_) << "\n";
        o << "    n: " << fmt_f32(m.n_) << "\n";
        o << "    glow_dist: " << fmt_f32(m.glow_dist_) << "\n";
        emit_color(o, "    ", "frac", m.frac_);
        o << "    pattern: " << enum_name(m.pattern_, PATTERNS) << "\n";
        o << "    pattern_scale: " << fmt_f32(m.pattern_scale_) << "\n";
        o << "    pattern_angle_scale: " << fmt_f32(m.pattern_angle_scale_) << "\n";
        o << "    texture_name: " << fmt_str(m.texture_name_) << "\n";
        o << "    texture_filter: " << enum_name(m.texture_filter_, FILTERS) << "\n";
    }
    if (objects_.empty()) o << "objects: []\n";
    else o << "objects:\n";
    for (const RenderObject &ob : objects_) {
        if (ob.kind == RenderObject::Sphere) {
            o << "  - Sphere:\n";
            o << "      material: " << fmt_str(ob.material->name_) << "\n";
            o << "      r: " << fmt_f32(ob.r) << "\n";
            emit_vec3(o, "      ", "org", ob.org);
        } else {
            o << "  - Floor:\n";
            o << "      material: " << fmt_str(ob.material->name_) << "\n";
            emit_vec3(o, "      ", "org", ob.org);
            emit_vec3(o, "      ", "face_normal", ob.face_normal);
        }
        o << "      uvmap: " << enum_name(ob.uvmap_, UVMAPS) << "\n";
    }
    return o.str();
}

// RenderEnv::deserialize, render.rs:762-799
void RenderEnv::deserialize(const std::string &s) {
    Node root;
    std::map<std::string, MaterialRef> mm;
    Camera cam;
    std::vector<CameraKeyframe> motion;
    int max_refl, max_refr;
    const Node *objs;
    try {
        root = parse_yaml(s);
        auto camera_of = [](const Node &n) { return Camera(as_vec3(field(n, "position")), as_vec3(field(n, "pyr"))); };
        cam = camera_of(field(root, "camera"));
        const Node &cm = field(root, "camera_motion");
        if (cm.type != Node::Seq && cm.type != Node::Null) throw ParseError("camera_motion must be a sequence");
        for (const Node &k : cm.seq) {
            CameraKeyframe kf;
            kf.camera = camera_of(field(k, "camera"));
            kf.velocity = as_vec3(field(k, "velocity"));
            const Node &t = field(k, "camera_target");
            kf.has_target = t.type != Node::Null;
            if (kf.has_target) kf.camera_target = as_vec3(t);
            kf.duration = as_f32(field(k, "duration"));
            motion.push_back(kf);
        }
        max_refl = as_i32(field(root, "max_reflections"));
        max_refr = as_i32(field(root, "max_refractions"));
        const Node &mats = field(root, "materials");
        if (mats.type != Node::Map && mats.type != Node::Null) throw ParseError("materials must be a mapping");
        for (auto &kv : mats.map) {  // RenderMaterial::deserialize, render.rs:201-218
            const Node &m = kv.second;
            auto mat = std::make_shared<RenderMaterial>(as_str(field(m, "name")), as_color(field(m, "diffuse")), as_color(field(m, "specular")),
                                                        as_i32(field(m, "pn")), as_f32(field(m, "t")), as_f32(field(m, "n")));
            mat->glow_dist_ = as_f32(field(m, "glow_dist"));
            mat->frac_ = as_color(field(m, "frac"));
            mat->pattern_ = as_enum(field(m, "pattern"), PATTERNS);
            mat->pattern_scale_ = as_f32(field(m, "pattern_scale"));
            mat->pattern_angle_scale_ = as_f32(field(m, "pattern_angle_scale"));
            mat->texture_name_ = as_str(field(m, "texture_name"));
            mat->texture_filter_ = as_enum(field(m, "texture_filter"), FILTERS);
            if (!mat->texture_name_.empty()) mat->texture_ = load_image_rgb8(mat->texture_name_);  // image::open(..).ok()
            mm[kv.first] = mat;
        }
        objs = &field(root, "objects");
        if (objs->type != Node::Seq && objs->type != Node::Null) throw ParseError("objects must be a sequence");
        for (const Node &o : objs->seq) {  // validate shapes before mutating self
            if (o.type != Node::Map || o.map.size() != 1 || (o.map[0].first != "Sphere" && o.map[0].first != "Floor"))
                throw ParseError("unknown object variant");
            const Node &b = o.map[0].second;
            as_str(field(b, "material"));
            as_vec3(field(b, "org"));
            as_enum(field(b, "uvmap"), UVMAPS);
            if (o.map[0].first == "Sphere") as_f32(field(b, "r"));
            else as_vec3(field(b, "face_normal"));
        }
    } catch (const ParseError &) {
        throw DeserializeError("serde_yaml::Error");  // From<serde_yaml::Error>, render.rs:360-366
    }
    camera = cam;
    camera_motion = motion;
    max_reflections = max_refl;
    max_refractions = max_refr;
    materials_ = mm;
    objects_.clear();
    invalidate();
    for (const Node &o : objs->seq) {
        const std::string &kind = o.map[0].first;
        const Node &b = o.map[0].second;
        const std::string mname = as_str(field(b, "material"));
        auto it = materials_.find(mname);
        if (it == materials_.end())
            throw DeserializeError((kind == "Sphere" ? "RenderSphere couldn't find material " : "RenderFloor couldn't find material ") + mname);
        if (kind == "Sphere")
            objects_.push_back(RenderSphere::make(it->second, as_f32(field(b, "r")), as_vec3(field(b, "org"))).uvmap(as_enum(field(b, "uvmap"), UVMAPS)));
        else
            objects_.push_back(RenderFloor::new_raw(it->second, as_vec3(field(b, "org")), as_vec3(field(b, "face_normal"))).uvmap(as_enum(field(b, "uvmap"), UVMAPS)));
    }
}

}  // namespace rr
