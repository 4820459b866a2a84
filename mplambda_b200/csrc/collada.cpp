// COLLADA (.dae) -> triangle soup, host side, no dependencies.
//
// Replaces MeshLoad<Mesh>::load (reference include/mpl/demo/load_mesh.hpp:232-293) for the
// subset of COLLADA the reference's resources use: <triangles>/<polylist>/<lines> primitive
// groups, nested <node><matrix>, <up_axis>, <unit>.  The emitted soup is what visitTriangles
// (load_mesh.hpp:50-89) hands to addTriangle: node transforms accumulated root->leaf in double
// from float32 matrices and float32 vertices (load_mesh.hpp:60,75-78), faces with <3 indices
// skipped, polygons fanned around corner 0.  The optional centre shift subtracts the mean of all
// vertices of all meshes (load_mesh.hpp:263-273), counted after the importer's
// JoinIdenticalVertices step, i.e. unique (position, normal, texcoord) tuples per primitive group.
#include "mplb_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>

namespace mplb {
namespace {

struct XmlNode {
    std::string tag;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::string text;
    std::vector<std::unique_ptr<XmlNode>> kids;

    const std::string *attr(const char *k) const {
        for (auto &a : attrs)
            if (a.first == k) return &a.second;
        return nullptr;
    }
    const XmlNode *child(const char *t) const {
        for (auto &c : kids)
            if (c->tag == t) return c.get();
        return nullptr;
    }
};

// Small non-validating XML reader: elements, attributes, character data.  Comments, processing
// instructions, CDATA and DOCTYPE are skipped.  Namespace prefixes are kept verbatim.
class XmlReader {
public:
    explicit XmlReader(const std::string &s) : s_(s) {}

    std::unique_ptr<XmlNode> parse() {
        std::unique_ptr<XmlNode> root;
        std::vector<XmlNode *> stack;
        while (p_ < s_.size()) {
            if (s_[p_] != '<') {
                size_t e = s_.find('<', p_);
                if (e == std::string::npos) e = s_.size();
                if (!stack.empty()) stack.back()->text.append(s_, p_, e - p_);
                p_ = e;
                continue;
            }
            if (s_.compare(p_, 4, "<!--") == 0) { skipTo("-->"); continue; }
            if (s_.compare(p_, 2, "<?") == 0) { skipTo("?>"); continue; }
            if (s_.compare(p_, 9, "<![CDATA[") == 0) { skipTo("]]>"); continue; }
            if (s_.compare(p_, 2, "<!") == 0) { skipTo(">"); continue; }
            if (s_.compare(p_, 2, "</") == 0) {
                skipTo(">");
                if (stack.empty()) throw std::runtime_error("unbalanced closing tag");
                stack.pop_back();
                continue;
            }
            ++p_;
            auto node = std::make_unique<XmlNode>();
            node->tag = name();
            bool selfClose = false;
            for (;;) {
                ws();
                if (p_ >= s_.size()) throw std::runtime_error("unterminated tag");
                if (s_[p_] == '/') { selfClose = true; ++p_; continue; }
                if (s_[p_] == '>') { ++p_; break; }
                std::string k = name();
                ws();
                if (p_ < s_.size() && s_[p_] == '=') {
                    ++p_;
                    ws();
                    char q = s_[p_++];
                    size_t e = s_.find(q, p_);
                    if (e == std::string::npos) throw std::runtime_error("unterminated attribute");
                    node->attrs.emplace_back(k, s_.substr(p_, e - p_));
                    p_ = e + 1;
                } else {
                    node->attrs.emplace_back(k, "");
                }
            }
            XmlNode *raw = node.get();
            if (stack.empty()) {
                if (root) throw std::runtime_error("multiple root elements");
                root = std::move(node);
            } else {
                stack.back()->kids.push_back(std::move(node));
            }
            if (!selfClose) stack.push_back(raw);
        }
        if (!root) throw std::runtime_error("no root element");
        return root;
    }

private:
    void ws() { while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) ++p_; }
    std::string name() {
        size_t b = p_;
        while (p_ < s_.size() && !std::isspace((unsigned char)s_[p_]) && s_[p_] != '=' &&
               s_[p_] != '>' && s_[p_] != '/')
            ++p_;
        return s_.substr(b, p_ - b);
    }
    void skipTo(const char *end) {
        size_t e = s_.find(end, p_);
        p_ = (e == std::string::npos) ? s_.size() : e + std::strlen(end);
    }
    const std::string &s_;
    size_t p_ = 0;
};

void parseFloats(const std::string &t, std::vector<float> &out) {
    const char *p = t.c_str();
    char *e;
    for (;;) {
        float v = std::strtof(p, &e);
        if (e == p) break;
        out.push_back(v);
        p = e;
    }
}
void parseInts(const std::string &t, std::vector<long> &out) {
    const char *p = t.c_str();
    char *e;
    for (;;) {
        long v = std::strtol(p, &e, 10);
        if (e == p) break;
        out.push_back(v);
        p = e;
    }
}

struct Source { std::vector<float> data; int stride = 1; };
struct Input { std::string semantic, source; int offset = 0; };
struct Group {
    int kind = 0;  // 3 = triangles, 2 = lines, 0 = polylist
    std::vector<Input> inputs;
    std::vector<long> p, vcount;
    int stride = 1;
};
struct Geometry {
    std::map<std::string, Source> sources;
    std::map<std::string, std::vector<Input>> vertices;
    std::vector<Group> groups;
};

struct Mat4 {
    double m[4][4];
    static Mat4 identity() {
        Mat4 r{};
        for (int i = 0; i < 4; ++i) r.m[i][i] = 1.0;
        return r;
    }
    Mat4 operator*(const Mat4 &o) const {
        Mat4 r{};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double s = 0;
                for (int k = 0; k < 4; ++k) s += m[i][k] * o.m[k][j];
                r.m[i][j] = s;
            }
        return r;
    }
    void apply(const float *v, double *o) const {
        for (int i = 0; i < 3; ++i)
            o[i] = m[i][0] * (double)v[0] + m[i][1] * (double)v[1] + m[i][2] * (double)v[2] + m[i][3];
    }
};

Geometry parseGeometry(const XmlNode &mesh) {
    Geometry g;
    for (auto &c : mesh.kids) {
        if (c->tag == "source") {
            const std::string *id = c->attr("id");
            const XmlNode *fa = c->child("float_array");
            if (!id || !fa) continue;
            Source s;
            parseFloats(fa->text, s.data);
            if (const XmlNode *tc = c->child("technique_common"))
                if (const XmlNode *acc = tc->child("accessor"))
                    if (const std::string *st = acc->attr("stride")) s.stride = std::atoi(st->c_str());
            g.sources["#" + *id] = std::move(s);
        } else if (c->tag == "vertices") {
            const std::string *id = c->attr("id");
            if (!id) continue;
            std::vector<Input> ins;
            for (auto &i : c->kids)
                if (i->tag == "input") {
                    Input in;
                    if (auto *s = i->attr("semantic")) in.semantic = *s;
                    if (auto *s = i->attr("source")) in.source = *s;
                    ins.push_back(in);
                }
            g.vertices["#" + *id] = ins;
        } else if (c->tag == "triangles" || c->tag == "lines" || c->tag == "polylist") {
            Group gr;
            gr.kind = c->tag == "triangles" ? 3 : (c->tag == "lines" ? 2 : 0);
            int maxOff = 0;
            for (auto &i : c->kids) {
                if (i->tag == "input") {
                    Input in;
                    if (auto *s = i->attr("semantic")) in.semantic = *s;
                    if (auto *s = i->attr("source")) in.source = *s;
                    if (auto *s = i->attr("offset")) in.offset = std::atoi(s->c_str());
                    maxOff = std::max(maxOff, in.offset);
                    gr.inputs.push_back(in);
                } else if (i->tag == "p") {
                    parseInts(i->text, gr.p);
                } else if (i->tag == "vcount") {
                    parseInts(i->text, gr.vcount);
                }
            }
            gr.stride = maxOff + 1;
            g.groups.push_back(std::move(gr));
        } else if (c->tag == "polygons" || c->tag == "tristrips" || c->tag == "trifans") {
            throw std::runtime_error("unsupported COLLADA primitive <" + c->tag + ">");
        }
    }
    return g;
}

struct Loader {
    std::map<std::string, Geometry> geoms;
    std::vector<double> tris;       // 9 per triangle
    double vsum[3] = {0, 0, 0};
    size_t vcount = 0;

    void emitGroup(const Geometry &g, const Group &gr, const Mat4 &tf) {
        size_t ncorner = gr.p.size() / gr.stride;
        // resolve the per-corner attribute streams that make up the importer's vertex identity
        struct Stream { const Source *src; int offset; bool position; };
        std::vector<Stream> streams;
        for (auto &in : gr.inputs) {
            if (in.semantic == "VERTEX") {
                auto v = g.vertices.find(in.source);
                if (v == g.vertices.end()) throw std::runtime_error("missing <vertices> " + in.source);
                for (auto &vi : v->second) {
                    auto s = g.sources.find(vi.source);
                    if (s == g.sources.end()) throw std::runtime_error("missing <source> " + vi.source);
                    streams.push_back({&s->second, in.offset, vi.semantic == "POSITION"});
                }
            } else if (in.semantic == "NORMAL" || in.semantic == "TEXCOORD") {
                auto s = g.sources.find(in.source);
                if (s == g.sources.end()) throw std::runtime_error("missing <source> " + in.source);
                streams.push_back({&s->second, in.offset, false});
            }
        }
        const Stream *pos = nullptr;
        for (auto &s : streams)
            if (s.position) pos = &s;
        if (!pos) throw std::runtime_error("primitive group without POSITION");
        if (pos->src->stride < 3) throw std::runtime_error("POSITION stride < 3");

        std::vector<double> wpos(3 * ncorner);
        std::vector<std::vector<float>> keys(ncorner);
        for (size_t c = 0; c < ncorner; ++c) {
            for (auto &s : streams) {
                long idx = gr.p[c * gr.stride + s.offset];
                size_t base = (size_t)idx * s.src->stride;
                if (idx < 0 || base + s.src->stride > s.src->data.size())
                    throw std::runtime_error("index out of range in <p>");
                for (int k = 0; k < s.src->stride; ++k) keys[c].push_back(s.src->data[base + k]);
                if (s.position) tf.apply(&s.src->data[base], &wpos[3 * c]);
            }
        }
        // centroid contribution: unique vertex tuples of this group
        {
            std::vector<size_t> order(ncorner);
            for (size_t i = 0; i < ncorner; ++i) order[i] = i;
            std::sort(order.begin(), order.end(),
                      [&](size_t a, size_t b) { return keys[a] < keys[b]; });
            for (size_t i = 0; i < ncorner; ++i) {
                if (i > 0 && keys[order[i]] == keys[order[i - 1]]) continue;
                for (int k = 0; k < 3; ++k) vsum[k] += wpos[3 * order[i] + k];
                ++vcount;
            }
        }
        // faces
        size_t start = 0, nfaces;
        if (gr.kind == 0) nfaces = gr.vcount.size();
        else nfaces = ncorner / gr.kind;
        for (size_t f = 0; f < nfaces; ++f) {
            size_t n = gr.kind == 0 ? (size_t)gr.vcount[f] : (size_t)gr.kind;
            if (start + n > ncorner) throw std::runtime_error("<vcount> exceeds <p>");
            for (size_t k = 2; k < n; ++k) {
                const size_t idx[3] = {start, start + k - 1, start + k};
                for (size_t c : idx)
                    for (int d = 0; d < 3; ++d) tris.push_back(wpos[3 * c + d]);
            }
            start += n;
        }
    }

    void visit(const XmlNode &node, Mat4 tf) {
        for (auto &c : node.kids) {
            if (c->tag == "matrix") {
                std::vector<float> v;
                parseFloats(c->text, v);
                if (v.size() != 16) throw std::runtime_error("<matrix> needs 16 values");
                Mat4 m;
                for (int i = 0; i < 16; ++i) m.m[i / 4][i % 4] = (double)v[i];
                tf = tf * m;
            } else if (c->tag == "translate" || c->tag == "rotate" || c->tag == "scale" ||
                       c->tag == "lookat" || c->tag == "skew") {
                throw std::runtime_error("unsupported COLLADA transform <" + c->tag + ">");
            }
        }
        for (auto &c : node.kids)
            if (c->tag == "instance_geometry") {
                const std::string *url = c->attr("url");
                if (!url) continue;
                auto g = geoms.find(*url);
                if (g == geoms.end()) continue;
                for (auto &gr : g->second.groups) emitGroup(g->second, gr, tf);
            }
        for (auto &c : node.kids)
            if (c->tag == "node") visit(*c, tf);
    }
};

}  // namespace

TriangleSoup loadCollada(const std::string &path, bool shiftToCenter, bool identityRootTransform) {
    std::string text;
    {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::invalid_argument("could not load mesh file '" + path + "'");
        char buf[1 << 16];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
        std::fclose(f);
    }
    std::unique_ptr<XmlNode> root;
    try {
        root = XmlReader(text).parse();
    } catch (const std::exception &e) {
        throw std::invalid_argument("could not load mesh file '" + path + "': " + e.what());
    }
    if (root->tag != "COLLADA")
        throw std::invalid_argument("could not load mesh file '" + path + "': not a COLLADA document");

    Loader L;
    try {
        if (const XmlNode *lg = root->child("library_geometries"))
            for (auto &g : lg->kids) {
                const std::string *id = g->attr("id");
                const XmlNode *mesh = g->child("mesh");
                if (g->tag == "geometry" && id && mesh) L.geoms["#" + *id] = parseGeometry(*mesh);
            }
        Mat4 rootTf = Mat4::identity();
        if (!identityRootTransform) {
            if (const XmlNode *asset = root->child("asset")) {
                if (const XmlNode *up = asset->child("up_axis")) {
                    std::string u = up->text;
                    u.erase(std::remove_if(u.begin(), u.end(), [](char c) { return std::isspace((unsigned char)c); }), u.end());
                    if (u == "Z_UP") {
                        Mat4 m{};  // (x,y,z) -> (x,z,-y)
                        m.m[0][0] = 1; m.m[1][2] = 1; m.m[2][1] = -1; m.m[3][3] = 1;
                        rootTf = rootTf * m;
                    } else if (u == "X_UP") {
                        Mat4 m{};
                        m.m[0][1] = -1; m.m[1][0] = 1; m.m[2][2] = 1; m.m[3][3] = 1;
                        rootTf = rootTf * m;
                    }
                }
                if (const XmlNode *unit = asset->child("unit"))
                    if (const std::string *meter = unit->attr("meter")) {
                        double s = (double)std::strtof(meter->c_str(), nullptr);
                        Mat4 m = Mat4::identity();
                        m.m[0][0] = m.m[1][1] = m.m[2][2] = s;
                        rootTf = rootTf * m;
                    }
            }
        }
        std::string sceneUrl;
        if (const XmlNode *sc = root->child("scene"))
            if (const XmlNode *ivs = sc->child("instance_visual_scene"))
                if (const std::string *u = ivs->attr("url")) sceneUrl = *u;
        const XmlNode *vscene = nullptr;
        if (const XmlNode *lvs = root->child("library_visual_scenes"))
            for (auto &v : lvs->kids) {
                const std::string *id = v->attr("id");
                if (v->tag == "visual_scene" && (!vscene || (id && "#" + *id == sceneUrl))) vscene = v.get();
            }
        if (!vscene) throw std::runtime_error("no visual_scene");
        for (auto &n : vscene->kids)
            if (n->tag == "node") L.visit(*n, rootTf);
    } catch (const std::invalid_argument &) {
        throw;
    } catch (const std::exception &e) {
        throw std::invalid_argument("could not load mesh file '" + path + "': " + e.what());
    }
    if (L.tris.empty())
        throw std::invalid_argument("mesh file '" + path + "' does not contain meshes");

    TriangleSoup out;
    out.tris = std::move(L.tris);
    if (shiftToCenter && L.vcount) {
        for (int k = 0; k < 3; ++k) out.center[k] = L.vsum[k] / (double)L.vcount;
        for (size_t i = 0; i < out.tris.size(); ++i) out.tris[i] -= out.center[i % 3];
    }
    return out;
}

}  // namespace mplb
