// mjcf_compile.cpp — MJCF-subset model compiler (host side, not on the hot path).
//
// Replaces the role of `mj_loadXML` (/root/reference/cmd/basic.cpp:123,
// /root/reference/tst/test_derivatives.cpp:34): turns the MJCF the reference ships
// (/root/reference/res/inverted_pendulum.xml, hopper.xml, humanoid.xml) into the flat
// `ilqg_model` tables of include/ilqg_model.h.  Supported subset: <compiler angle coordinate
// inertiafromgeom>, one class-less <default> (joint/geom/motor), <option>, a <worldbody> tree of
// <body>/<joint>/<freejoint>/<geom> with plane/sphere/capsule geoms, and <actuator><motor>.
// Anything else that would change the dynamics is rejected with an error string.
//
// Constants that MuJoCo derives at compile time (body inertias from geoms at density 1000,
// dof_invweight0 / body_invweight0 / meaninertia at qpos0) are computed here with a dense
// Jacobian formulation  M = sum_b m Jp'Jp + Jr' I Jr  — deliberately a different algorithm from
// the CRBA used by the oracle and the kernels, so tests can cross-check the three.
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ilqg_b200.h"

namespace {

// ---------------------------------------------------------------- tiny XML reader
struct Elem {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<Elem>> kids;
    const std::string* get(const char* k) const {
        for (auto& a : attrs)
            if (a.first == k) return &a.second;
        return nullptr;
    }
    const Elem* child(const char* n) const {
        for (auto& c : kids)
            if (c->name == n) return c.get();
        return nullptr;
    }
};

struct XmlReader {
    const std::string& s;
    size_t p = 0;
    std::string err;
    explicit XmlReader(const std::string& src) : s(src) {}
    void ws() {
        while (p < s.size() && isspace((unsigned char)s[p])) p++;
    }
    bool skipMisc() {  // comments, processing instructions, text
        for (;;) {
            while (p < s.size() && s[p] != '<') p++;
            if (p >= s.size()) return false;
            if (s.compare(p, 4, "<!--") == 0) {
                size_t e = s.find("-->", p + 4);
                if (e == std::string::npos) { err = "unterminated comment"; return false; }
                p = e + 3;
            } else if (s.compare(p, 2, "<?") == 0) {
                size_t e = s.find("?>", p + 2);
                if (e == std::string::npos) { err = "unterminated <?"; return false; }
                p = e + 2;
            } else if (s.compare(p, 2, "<!") == 0) {
                size_t e = s.find('>', p);
                if (e == std::string::npos) { err = "unterminated <!"; return false; }
                p = e + 1;
            } else
                return true;
        }
    }
    std::string ident() {
        size_t b = p;
        while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '_' || s[p] == '-' || s[p] == ':' || s[p] == '.')) p++;
        return s.substr(b, p - b);
    }
    std::unique_ptr<Elem> element() {
        // precondition: s[p]=='<' and it is an opening tag
        p++;
        auto e = std::make_unique<Elem>();
        e->name = ident();
        if (e->name.empty()) { err = "bad tag"; return nullptr; }
        for (;;) {
            ws();
            if (p >= s.size()) { err = "eof in tag"; return nullptr; }
            if (s[p] == '/') {
                if (p + 1 < s.size() && s[p + 1] == '>') { p += 2; return e; }
                err = "bad '/'"; return nullptr;
            }
            if (s[p] == '>') { p++; break; }
            std::string k = ident();
            ws();
            if (k.empty() || p >= s.size() || s[p] != '=') { err = "bad attribute in <" + e->name + ">"; return nullptr; }
            p++;
            ws();
            char q = s[p];
            if (q != '"' && q != '\'') { err = "unquoted attribute"; return nullptr; }
            size_t c = s.find(q, p + 1);
            if (c == std::string::npos) { err = "unterminated attribute"; return nullptr; }
            e->attrs.emplace_back(k, s.substr(p + 1, c - p - 1));
            p = c + 1;
        }
        for (;;) {  // children until </name>
            if (!skipMisc()) { if (err.empty()) err = "eof inside <" + e->name + ">"; return nullptr; }
            if (s.compare(p, 2, "</") == 0) {
                p += 2;
                std::string n = ident();
                ws();
                if (n != e->name || p >= s.size() || s[p] != '>') { err = "mismatched </" + n + ">"; return nullptr; }
                p++;
                return e;
            }
            auto k = element();
            if (!k) return nullptr;
            e->kids.push_back(std::move(k));
        }
    }
    std::unique_ptr<Elem> parse() {
        if (!skipMisc()) { if (err.empty()) err = "no root element"; return nullptr; }
        return element();
    }
};

// ---------------------------------------------------------------- small math
struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
struct Q4 { double w, x, y, z; };
inline Q4 qmul(Q4 a, Q4 b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
inline Q4 qconj(Q4 a) { return {a.w, -a.x, -a.y, -a.z}; }
inline Q4 qnormalize(Q4 a) {
    double n = std::sqrt(a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z);
    if (n < 1e-15) return {1, 0, 0, 0};
    return {a.w / n, a.x / n, a.y / n, a.z / n};
}
struct M3 { double m[3][3]; };
inline M3 q2m(Q4 q) {
    M3 r;
    double w = q.w, x = q.x, y = q.y, z = q.z;
    r.m[0][0] = w * w + x * x - y * y - z * z; r.m[0][1] = 2 * (x * y - w * z); r.m[0][2] = 2 * (x * z + w * y);
    r.m[1][0] = 2 * (x * y + w * z); r.m[1][1] = w * w - x * x + y * y - z * z; r.m[1][2] = 2 * (y * z - w * x);
    r.m[2][0] = 2 * (x * z - w * y); r.m[2][1] = 2 * (y * z + w * x); r.m[2][2] = w * w - x * x - y * y + z * z;
    return r;
}
inline V3 mulv(const M3& R, V3 v) {
    return {R.m[0][0] * v.x + R.m[0][1] * v.y + R.m[0][2] * v.z, R.m[1][0] * v.x + R.m[1][1] * v.y + R.m[1][2] * v.z,
            R.m[2][0] * v.x + R.m[2][1] * v.y + R.m[2][2] * v.z};
}
inline V3 multv(const M3& R, V3 v) {  // R^T v
    return {R.m[0][0] * v.x + R.m[1][0] * v.y + R.m[2][0] * v.z, R.m[0][1] * v.x + R.m[1][1] * v.y + R.m[2][1] * v.z,
            R.m[0][2] * v.x + R.m[1][2] * v.y + R.m[2][2] * v.z};
}
inline M3 mmul(const M3& A, const M3& B) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
inline M3 mtrans(const M3& A) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C.m[i][j] = A.m[j][i];
    return C;
}
// quaternion taking +z onto unit vector v
inline Q4 z2quat(V3 v) {
    V3 ax = cross({0, 0, 1}, v);
    double s = norm(ax);
    if (s < 1e-10) ax = {1, 0, 0};
    else ax = (1.0 / s) * ax;
    double ang = std::atan2(s, v.z);
    return qnormalize({std::cos(ang / 2), ax.x * std::sin(ang / 2), ax.y * std::sin(ang / 2), ax.z * std::sin(ang / 2)});
}

// ---------------------------------------------------------------- attribute helpers
struct Ctx {
    std::string err;
    bool degree = true, global = false;
    const Elem* defJoint = nullptr;
    const Elem* defGeom = nullptr;
    const Elem* defMotor = nullptr;
    bool fail(const std::string& m) { if (err.empty()) err = m; return false; }
};

const std::string* attr(const Elem* e, const Elem* def, const char* k) {
    if (const std::string* v = e->get(k)) return v;
    if (def) return def->get(k);
    return nullptr;
}
// read up to n numbers; returns count read. strtod prefix semantics ("0.13/2" -> 0.13).
int nums(const std::string* s, double* out, int n) {
    if (!s) return 0;
    const char* c = s->c_str();
    int k = 0;
    while (*c && k < n) {
        while (*c && isspace((unsigned char)*c)) c++;
        if (!*c) break;
        char* e;
        double v = strtod(c, &e);
        if (e == c) break;
        out[k++] = v;
        c = e;
        while (*c && !isspace((unsigned char)*c)) c++;  // drop trailing junk of this token
    }
    return k;
}
bool truthy(const std::string* s, bool dflt) {
    if (!s) return dflt;
    return *s == "true";
}

struct BodyTmp {
    V3 gpos; Q4 gquat;  // global frame at qpos0
};

}  // namespace

// ---------------------------------------------------------------- the compiler
namespace {

struct Compiler {
    Ctx cx;
    ilqg_model* m;
    std::vector<BodyTmp> bt;
    std::map<std::string, int> jntByName;

    bool addGeom(const Elem* e, int body) {
        if (m->ngeom >= ILQG_MAXGEOM) return cx.fail("too many geoms");
        int g = m->ngeom++;
        const Elem* D = cx.defGeom;
        const std::string* ty = attr(e, D, "type");
        int type = ILQG_GEOM_SPHERE;
        if (ty) {
            if (*ty == "plane") type = ILQG_GEOM_PLANE;
            else if (*ty == "sphere") type = ILQG_GEOM_SPHERE;
            else if (*ty == "capsule") type = ILQG_GEOM_CAPSULE;
            else return cx.fail("unsupported geom type '" + *ty + "'");
        }
        m->geom_type[g] = type;
        m->geom_bodyid[g] = body;
        double sz[3] = {0, 0, 0};
        nums(attr(e, D, "size"), sz, 3);
        V3 pos = {0, 0, 0};
        Q4 quat = {1, 0, 0, 0};
        double t[6];
        if (nums(e->get("pos"), t, 3) == 3) pos = {t[0], t[1], t[2]};
        if (nums(e->get("quat"), t, 4) == 4) quat = qnormalize({t[0], t[1], t[2], t[3]});
        if (e->get("euler") || e->get("axisangle") || e->get("zaxis") || e->get("xyaxes")) return cx.fail("geom orientation spec not supported");
        if (nums(e->get("fromto"), t, 6) == 6) {
            if (type != ILQG_GEOM_CAPSULE) return cx.fail("fromto on non-capsule");
            V3 a = {t[0], t[1], t[2]}, b = {t[3], t[4], t[5]};
            V3 d = a - b;  // MuJoCo aligns +z with (from - to)
            double len = norm(d);
            if (len < 1e-12) return cx.fail("degenerate fromto");
            pos = 0.5 * (a + b);
            quat = z2quat((1.0 / len) * d);
            sz[1] = len / 2;
        }
        // geom frame is relative to the body frame, or global under coordinate="global"
        if (cx.global) {
            M3 Rb = q2m(bt[body].gquat);
            pos = multv(Rb, pos - bt[body].gpos);
            quat = qnormalize(qmul(qconj(bt[body].gquat), quat));
        }
        m->geom_size[g][0] = sz[0]; m->geom_size[g][1] = sz[1]; m->geom_size[g][2] = sz[2];
        m->geom_pos[g][0] = pos.x; m->geom_pos[g][1] = pos.y; m->geom_pos[g][2] = pos.z;
        m->geom_quat[g][0] = quat.w; m->geom_quat[g][1] = quat.x; m->geom_quat[g][2] = quat.y; m->geom_quat[g][3] = quat.z;
        double v[5];
        m->geom_contype[g] = nums(attr(e, D, "contype"), v, 1) ? (int)v[0] : 1;
        m->geom_conaffinity[g] = nums(attr(e, D, "conaffinity"), v, 1) ? (int)v[0] : 1;
        m->geom_condim[g] = nums(attr(e, D, "condim"), v, 1) ? (int)v[0] : 3;
        if (m->geom_condim[g] != 1 && m->geom_condim[g] != 3) return cx.fail("only condim 1 and 3 are supported");
        double fr[3] = {1, 0.005, 0.0001};
        nums(attr(e, D, "friction"), fr, 3);
        for (int i = 0; i < 3; i++) m->geom_friction[g][i] = fr[i];
        m->geom_margin[g] = nums(attr(e, D, "margin"), v, 1) ? v[0] : 0.0;
        m->geom_gap[g] = nums(attr(e, D, "gap"), v, 1) ? v[0] : 0.0;
        double sr[2] = {0.02, 1};
        nums(attr(e, D, "solref"), sr, 2);
        m->geom_solref[g][0] = sr[0]; m->geom_solref[g][1] = sr[1];
        double si[5] = {0.9, 0.95, 0.001, 0.5, 2};
        nums(attr(e, D, "solimp"), si, 5);
        for (int i = 0; i < 5; i++) m->geom_solimp[g][i] = si[i];
        m->geom_solmix[g] = nums(attr(e, D, "solmix"), v, 1) ? v[0] : 1.0;
        density[g] = nums(attr(e, D, "density"), v, 1) ? v[0] : 1000.0;
        if (attr(e, D, "mass")) return cx.fail("geom mass attribute not supported");
        return true;
    }
    double density[ILQG_MAXGEOM];

    bool addJoint(const Elem* e, int body, bool freejoint) {
        if (m->njnt >= ILQG_MAXJNT) return cx.fail("too many joints");
        int j = m->njnt++;
        const Elem* D = freejoint ? nullptr : cx.defJoint;
        int type = ILQG_JNT_HINGE;
        if (freejoint) type = ILQG_JNT_FREE;
        else if (const std::string* ty = attr(e, D, "type")) {
            if (*ty == "hinge") type = ILQG_JNT_HINGE;
            else if (*ty == "slide") type = ILQG_JNT_SLIDE;
            else if (*ty == "free") type = ILQG_JNT_FREE;
            else if (*ty == "ball") type = ILQG_JNT_BALL;
            else return cx.fail("unsupported joint type '" + *ty + "'");
        }
        if (const std::string* nm = e->get("name")) jntByName[*nm] = j;
        m->jnt_type[j] = type;
        m->jnt_bodyid[j] = body;
        m->jnt_qposadr[j] = m->nq;
        m->jnt_dofadr[j] = m->nv;
        int nq = type == ILQG_JNT_FREE ? 7 : (type == ILQG_JNT_BALL ? 4 : 1), nd = type == ILQG_JNT_FREE ? 6 : (type == ILQG_JNT_BALL ? 3 : 1);
        if (m->nq + nq > ILQG_MAXQ || m->nv + nd > ILQG_MAXV) return cx.fail("too many dofs");
        double t[5];
        V3 pos = {0, 0, 0}, axis = {0, 0, 1};
        if (nums(attr(e, D, "pos"), t, 3) == 3) pos = {t[0], t[1], t[2]};
        if (nums(attr(e, D, "axis"), t, 3) == 3) axis = {t[0], t[1], t[2]};
        double an = norm(axis);
        if (an < 1e-12) return cx.fail("zero joint axis");
        axis = (1.0 / an) * axis;
        if (cx.global && type != ILQG_JNT_FREE) {
            M3 Rb = q2m(bt[body].gquat);
            pos = multv(Rb, pos - bt[body].gpos);
            axis = multv(Rb, axis);
        }
        m->jnt_pos[j][0] = pos.x; m->jnt_pos[j][1] = pos.y; m->jnt_pos[j][2] = pos.z;
        m->jnt_axis[j][0] = axis.x; m->jnt_axis[j][1] = axis.y; m->jnt_axis[j][2] = axis.z;
        double ang = (cx.degree && type == ILQG_JNT_HINGE) ? M_PI / 180.0 : 1.0;
        double rg[2] = {0, 0};
        nums(attr(e, D, "range"), rg, 2);
        m->jnt_range[j][0] = rg[0] * ang; m->jnt_range[j][1] = rg[1] * ang;
        m->jnt_limited[j] = truthy(attr(e, D, "limited"), false) ? 1 : 0;
        if (type == ILQG_JNT_FREE) m->jnt_limited[j] = 0;
        if (type == ILQG_JNT_BALL && m->jnt_limited[j]) return cx.fail("limited ball joints are not supported");
        m->jnt_stiffness[j] = nums(attr(e, D, "stiffness"), t, 1) ? t[0] : 0.0;
        m->jnt_margin[j] = nums(attr(e, D, "margin"), t, 1) ? t[0] : 0.0;
        double sr[2] = {0.02, 1};
        nums(attr(e, D, "solreflimit"), sr, 2);
        m->jnt_solref[j][0] = sr[0]; m->jnt_solref[j][1] = sr[1];
        double si[5] = {0.9, 0.95, 0.001, 0.5, 2};
        nums(attr(e, D, "solimplimit"), si, 5);
        for (int i = 0; i < 5; i++) m->jnt_solimp[j][i] = si[i];
        double arm = nums(attr(e, D, "armature"), t, 1) ? t[0] : 0.0;
        double damp = nums(attr(e, D, "damping"), t, 1) ? t[0] : 0.0;
        if (attr(e, D, "frictionloss")) return cx.fail("frictionloss not supported");
        double ref = nums(attr(e, D, "ref"), t, 1) ? t[0] * ang : 0.0;
        double sref = nums(attr(e, D, "springref"), t, 1) ? t[0] * ang : 0.0;
        if (type == ILQG_JNT_FREE) {
            const BodyTmp& b = bt[body];
            double q0[7] = {b.gpos.x, b.gpos.y, b.gpos.z, b.gquat.w, b.gquat.x, b.gquat.y, b.gquat.z};
            for (int i = 0; i < 7; i++) m->qpos0[m->nq + i] = m->qpos_spring[m->nq + i] = q0[i];
            arm = 0; damp = 0;  // freejoint carries no defaults
        } else if (type == ILQG_JNT_BALL) {   // the joint's own rotation: identity at qpos0
            if (m->jnt_stiffness[j] != 0) return cx.fail("springs on ball joints are not supported");
            const double q0[4] = {1, 0, 0, 0};
            for (int i = 0; i < 4; i++) m->qpos0[m->nq + i] = m->qpos_spring[m->nq + i] = q0[i];
        } else {
            m->qpos0[m->nq] = ref;
            m->qpos_spring[m->nq] = sref;
        }
        for (int i = 0; i < nd; i++) {
            int d = m->nv + i;
            m->dof_bodyid[d] = body;
            m->dof_jntid[d] = j;
            m->dof_armature[d] = arm;
            m->dof_damping[d] = damp;
        }
        m->nq += nq;
        m->nv += nd;
        return true;
    }

    bool addBody(const Elem* e, int parent) {
        if (m->nbody >= ILQG_MAXBODY) return cx.fail("too many bodies");
        int b = m->nbody++;
        m->body_parentid[b] = parent;
        m->body_rootid[b] = parent == 0 ? b : m->body_rootid[parent];
        double t[4];
        V3 pos = {0, 0, 0};
        Q4 quat = {1, 0, 0, 0};
        if (nums(e->get("pos"), t, 3) == 3) pos = {t[0], t[1], t[2]};
        if (nums(e->get("quat"), t, 4) == 4) quat = qnormalize({t[0], t[1], t[2], t[3]});
        if (e->get("euler") || e->get("axisangle") || e->get("zaxis") || e->get("xyaxes")) return cx.fail("body orientation spec not supported");
        BodyTmp me;
        const BodyTmp& P = bt[parent];
        M3 Rp = q2m(P.gquat);
        if (cx.global) {
            me.gpos = pos; me.gquat = quat;
            pos = multv(Rp, pos - P.gpos);
            quat = qnormalize(qmul(qconj(P.gquat), quat));
        } else {
            me.gpos = P.gpos + mulv(Rp, pos);
            me.gquat = qnormalize(qmul(P.gquat, quat));
        }
        bt.push_back(me);
        m->body_pos[b][0] = pos.x; m->body_pos[b][1] = pos.y; m->body_pos[b][2] = pos.z;
        m->body_quat[b][0] = quat.w; m->body_quat[b][1] = quat.x; m->body_quat[b][2] = quat.y; m->body_quat[b][3] = quat.z;
        m->body_jntadr[b] = m->njnt;
        m->body_dofadr[b] = m->nv;
        for (auto& k : e->kids) {
            if (k->name == "joint") { if (!addJoint(k.get(), b, false)) return false; }
            else if (k->name == "freejoint") { if (!addJoint(k.get(), b, true)) return false; }
        }
        m->body_jntnum[b] = m->njnt - m->body_jntadr[b];
        m->body_dofnum[b] = m->nv - m->body_dofadr[b];
        for (int j = m->body_jntadr[b]; j < m->njnt; j++)   // (its motion axes are the body's own: MuJoCo takes them from the body frame)
            if (m->jnt_type[j] == ILQG_JNT_BALL && m->body_jntnum[b] != 1) return cx.fail("a ball joint must be the only joint of its body");
        for (auto& k : e->kids)
            if (k->name == "geom") { if (!addGeom(k.get(), b)) return false; }
        if (e->child("inertial")) return cx.fail("<inertial> not supported (inertiafromgeom only)");
        for (auto& k : e->kids) {
            if (k->name == "body") { if (!addBody(k.get(), b)) return false; }
            else if (k->name != "joint" && k->name != "freejoint" && k->name != "geom" && k->name != "site" && k->name != "light" &&
                     k->name != "camera")
                return cx.fail("unsupported element <" + k->name + "> in <body>");
        }
        return true;
    }

    // mass, centre of mass and inertia of every body from its geoms
    bool inertiaFromGeoms() {
        for (int b = 1; b < m->nbody; b++) {
            double mass = 0;
            V3 com = {0, 0, 0};
            struct GI { double mass; V3 pos; M3 I; };
            std::vector<GI> gs;
            for (int g = 0; g < m->ngeom; g++) {
                if (m->geom_bodyid[g] != b) continue;
                double r = m->geom_size[g][0], h = m->geom_size[g][1];
                GI gi;
                double I[3];
                if (m->geom_type[g] == ILQG_GEOM_SPHERE) {
                    gi.mass = density[g] * 4.0 / 3.0 * M_PI * r * r * r;
                    I[0] = I[1] = I[2] = 0.4 * gi.mass * r * r;
                } else if (m->geom_type[g] == ILQG_GEOM_CAPSULE) {
                    double H = 2 * h;
                    double vcyl = M_PI * r * r * H, vsph = 4.0 / 3.0 * M_PI * r * r * r;
                    gi.mass = density[g] * (vcyl + vsph);
                    double mc = density[g] * vcyl, ms = density[g] * vsph;
                    // cylinder + two hemispheres (hemisphere com sits 3r/8 beyond the cylinder end)
                    double Ixx = mc * (3 * r * r + H * H) / 12.0 + ms * (0.4 * r * r + H * H / 4.0 + 3.0 * H * r / 8.0);
                    I[0] = I[1] = Ixx;
                    I[2] = mc * r * r / 2.0 + 0.4 * ms * r * r;
                } else
                    continue;  // planes are massless
                gi.pos = {m->geom_pos[g][0], m->geom_pos[g][1], m->geom_pos[g][2]};
                M3 R = q2m({m->geom_quat[g][0], m->geom_quat[g][1], m->geom_quat[g][2], m->geom_quat[g][3]});
                M3 Dg = {{{I[0], 0, 0}, {0, I[1], 0}, {0, 0, I[2]}}};
                gi.I = mmul(mmul(R, Dg), mtrans(R));
                mass += gi.mass;
                com = com + gi.mass * gi.pos;
                gs.push_back(gi);
            }
            if (mass <= 0) return cx.fail("body without mass");
            com = (1.0 / mass) * com;
            M3 I = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
            for (auto& gi : gs) {
                V3 d = gi.pos - com;
                double dd = dot(d, d);
                double dv[3] = {d.x, d.y, d.z};
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) I.m[i][j] += gi.I.m[i][j] + gi.mass * ((i == j ? dd : 0.0) - dv[i] * dv[j]);
            }
            m->body_mass[b] = mass;
            m->body_ipos[b][0] = com.x; m->body_ipos[b][1] = com.y; m->body_ipos[b][2] = com.z;
            double* o = m->body_inertia[b];
            o[0] = I.m[0][0]; o[1] = I.m[1][1]; o[2] = I.m[2][2]; o[3] = I.m[0][1]; o[4] = I.m[0][2]; o[5] = I.m[1][2];
        }
        return true;
    }

    void dofParents() {
        for (int d = 0; d < m->nv; d++) {
            int b = m->dof_bodyid[d];
            if (d > m->body_dofadr[b]) { m->dof_parentid[d] = d - 1; continue; }
            int p = m->body_parentid[b];
            while (p > 0 && m->body_dofnum[p] == 0) p = m->body_parentid[p];
            m->dof_parentid[d] = p > 0 ? m->body_dofadr[p] + m->body_dofnum[p] - 1 : -1;
        }
    }

    bool isAncestorOrSelf(int anc, int b) const {
        while (b > 0) { if (b == anc) return true; b = m->body_parentid[b]; }
        return false;
    }

    // dof_invweight0, body_invweight0, meaninertia at qpos0 (body frames = bt[])
    bool constantsAtQpos0() {
        int nv = m->nv;
        if (nv == 0) return cx.fail("model has no dofs");
        std::vector<double> M(nv * nv, 0.0);
        // per body: 6 x nv Jacobian at the centre of mass (rows 0-2 translational, 3-5 rotational)
        std::vector<std::vector<double>> J(m->nbody, std::vector<double>(6 * nv, 0.0));
        for (int b = 1; b < m->nbody; b++) {
            M3 Rb = q2m(bt[b].gquat);
            V3 p = bt[b].gpos + mulv(Rb, {m->body_ipos[b][0], m->body_ipos[b][1], m->body_ipos[b][2]});
            for (int d = 0; d < nv; d++) {
                int db = m->dof_bodyid[d];
                if (!isAncestorOrSelf(db, b)) continue;
                int j = m->dof_jntid[d];
                M3 Rj = q2m(bt[db].gquat);
                V3 jp = {0, 0, 0}, jr = {0, 0, 0};
                if (m->jnt_type[j] == ILQG_JNT_FREE) {
                    int k = d - m->jnt_dofadr[j];
                    if (k < 3) { double e[3] = {0, 0, 0}; e[k] = 1; jp = {e[0], e[1], e[2]}; }
                    else { V3 ax = {Rj.m[0][k - 3], Rj.m[1][k - 3], Rj.m[2][k - 3]}; jr = ax; jp = cross(ax, p - bt[db].gpos); }
                } else if (m->jnt_type[j] == ILQG_JNT_BALL) {   // the body's k-th axis about the joint anchor
                    int k = d - m->jnt_dofadr[j];
                    V3 ax = {Rj.m[0][k], Rj.m[1][k], Rj.m[2][k]};
                    V3 anchor = bt[db].gpos + mulv(Rj, {m->jnt_pos[j][0], m->jnt_pos[j][1], m->jnt_pos[j][2]});
                    jr = ax; jp = cross(ax, p - anchor);
                } else {
                    V3 ax = mulv(Rj, {m->jnt_axis[j][0], m->jnt_axis[j][1], m->jnt_axis[j][2]});
                    V3 anchor = bt[db].gpos + mulv(Rj, {m->jnt_pos[j][0], m->jnt_pos[j][1], m->jnt_pos[j][2]});
                    if (m->jnt_type[j] == ILQG_JNT_SLIDE) jp = ax;
                    else { jr = ax; jp = cross(ax, p - anchor); }
                }
                double* Jb = J[b].data();
                Jb[0 * nv + d] = jp.x; Jb[1 * nv + d] = jp.y; Jb[2 * nv + d] = jp.z;
                Jb[3 * nv + d] = jr.x; Jb[4 * nv + d] = jr.y; Jb[5 * nv + d] = jr.z;
            }
            const double* in = m->body_inertia[b];
            M3 Ib = {{{in[0], in[3], in[4]}, {in[3], in[1], in[5]}, {in[4], in[5], in[2]}}};
            M3 Iw = mmul(mmul(Rb, Ib), mtrans(Rb));
            const double* Jb = J[b].data();
            for (int r = 0; r < nv; r++)
                for (int c = 0; c < nv; c++) {
                    double s = 0;
                    for (int k = 0; k < 3; k++) s += m->body_mass[b] * Jb[k * nv + r] * Jb[k * nv + c];
                    for (int k = 0; k < 3; k++)
                        for (int l = 0; l < 3; l++) s += Jb[(3 + k) * nv + r] * Iw.m[k][l] * Jb[(3 + l) * nv + c];
                    M[r * nv + c] += s;
                }
        }
        double tr = 0;
        for (int d = 0; d < nv; d++) { M[d * nv + d] += m->dof_armature[d]; tr += M[d * nv + d]; }
        m->meaninertia = tr / nv;
        // Minv via Cholesky
        std::vector<double> L(M);
        for (int i = 0; i < nv; i++) {
            for (int j = 0; j <= i; j++) {
                double s = L[i * nv + j];
                for (int k = 0; k < j; k++) s -= L[i * nv + k] * L[j * nv + k];
                if (i == j) { if (s <= 0) return cx.fail("mass matrix not positive definite at qpos0"); L[i * nv + i] = std::sqrt(s); }
                else L[i * nv + j] = s / L[j * nv + j];
            }
        }
        auto solve = [&](std::vector<double>& x) {
            for (int i = 0; i < nv; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[i * nv + k] * x[k]; x[i] = s / L[i * nv + i]; }
            for (int i = nv - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < nv; k++) s -= L[k * nv + i] * x[k]; x[i] = s / L[i * nv + i]; }
        };
        std::vector<double> Minv(nv * nv);
        for (int c = 0; c < nv; c++) {
            std::vector<double> e(nv, 0.0);
            e[c] = 1;
            solve(e);
            for (int r = 0; r < nv; r++) Minv[r * nv + c] = e[r];
        }
        for (int j = 0; j < m->njnt; j++) {
            int a = m->jnt_dofadr[j];
            if (m->jnt_type[j] == ILQG_JNT_FREE) {
                double t = (Minv[a * nv + a] + Minv[(a + 1) * nv + a + 1] + Minv[(a + 2) * nv + a + 2]) / 3;
                double r = (Minv[(a + 3) * nv + a + 3] + Minv[(a + 4) * nv + a + 4] + Minv[(a + 5) * nv + a + 5]) / 3;
                for (int k = 0; k < 3; k++) { m->dof_invweight0[a + k] = t; m->dof_invweight0[a + 3 + k] = r; }
            } else if (m->jnt_type[j] == ILQG_JNT_BALL) {   // one value for the joint's three dofs, as MuJoCo averages them
                double r = (Minv[a * nv + a] + Minv[(a + 1) * nv + a + 1] + Minv[(a + 2) * nv + a + 2]) / 3;
                for (int k = 0; k < 3; k++) m->dof_invweight0[a + k] = r;
            } else
                m->dof_invweight0[a] = Minv[a * nv + a];
        }
        for (int b = 1; b < m->nbody; b++) {
            const double* Jb = J[b].data();
            double acc[2] = {0, 0};
            for (int k = 0; k < 6; k++) {
                double s = 0;
                for (int r = 0; r < nv; r++)
                    for (int c = 0; c < nv; c++) s += Jb[k * nv + r] * Minv[r * nv + c] * Jb[k * nv + c];
                acc[k / 3] += s;
            }
            m->body_invweight0[b][0] = acc[0] / 3;
            m->body_invweight0[b][1] = acc[1] / 3;
        }
        return true;
    }

    bool makePairs() {
        std::vector<int> weld(m->nbody, 0);
        for (int b = 1; b < m->nbody; b++) weld[b] = m->body_jntnum[b] ? b : weld[m->body_parentid[b]];
        for (int a = 0; a < m->ngeom; a++)
            for (int c = a + 1; c < m->ngeom; c++) {
                int g1 = a, g2 = c;
                if (m->geom_type[g1] > m->geom_type[g2]) { g1 = c; g2 = a; }
                int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
                int w1 = weld[b1], w2 = weld[b2];
                if (w1 == w2) continue;
                if (w1 != 0 && w2 != 0 && (weld[m->body_parentid[w1]] == w2 || weld[m->body_parentid[w2]] == w1)) continue;
                if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
                if (m->geom_type[g1] == ILQG_GEOM_PLANE && m->geom_type[g2] == ILQG_GEOM_PLANE) continue;
                if (m->npair >= ILQG_MAXPAIR) return cx.fail("too many collision pairs");
                int p = m->npair++;
                m->pair_geom1[p] = g1;
                m->pair_geom2[p] = g2;
                m->pair_condim[p] = std::max(m->geom_condim[g1], m->geom_condim[g2]);
                double margin = std::max(m->geom_margin[g1], m->geom_margin[g2]);
                double gap = std::max(m->geom_gap[g1], m->geom_gap[g2]);
                if (gap != 0) return cx.fail("geom gap not supported");
                m->pair_margin[p] = margin - gap;
                m->pair_friction[p] = std::max(m->geom_friction[g1][0], m->geom_friction[g2][0]);
                double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
                double mix = (s1 + s2) > 0 ? s1 / (s1 + s2) : 0.5;
                if (m->geom_solref[g1][0] <= 0 || m->geom_solref[g2][0] <= 0) return cx.fail("direct (negative) solref not supported");
                for (int i = 0; i < 2; i++) m->pair_solref[p][i] = mix * m->geom_solref[g1][i] + (1 - mix) * m->geom_solref[g2][i];
                for (int i = 0; i < 5; i++) m->pair_solimp[p][i] = mix * m->geom_solimp[g1][i] + (1 - mix) * m->geom_solimp[g2][i];
            }
        return true;
    }

    bool run(const Elem* root) {
        memset(m, 0, sizeof(*m));
        m->magic = ILQG_MODEL_MAGIC;
        m->version = ILQG_MODEL_VERSION;
        if (root->name != "mujoco") return cx.fail("root element is not <mujoco>");
        if (const Elem* c = root->child("compiler")) {
            if (const std::string* a = c->get("angle")) cx.degree = (*a != "radian");
            if (const std::string* a = c->get("coordinate")) cx.global = (*a == "global");
            if (const std::string* a = c->get("inertiafromgeom")) if (*a == "false") return cx.fail("inertiafromgeom=false not supported");
        }
        if (const Elem* d = root->child("default")) {
            cx.defJoint = d->child("joint");
            cx.defGeom = d->child("geom");
            cx.defMotor = d->child("motor");
            if (d->child("default")) return cx.fail("nested default classes not supported");
        }
        m->timestep = 0.002;
        m->gravity[0] = 0; m->gravity[1] = 0; m->gravity[2] = -9.81;
        m->tolerance = 1e-8;
        m->ls_tolerance = 0.01;
        m->impratio = 1;
        m->integrator = ILQG_INT_EULER;
        m->iterations = 100;
        m->ls_iterations = 50;
        if (const Elem* o = root->child("option")) {
            double t[3];
            if (nums(o->get("timestep"), t, 1)) m->timestep = t[0];
            if (nums(o->get("gravity"), t, 3) == 3) { m->gravity[0] = t[0]; m->gravity[1] = t[1]; m->gravity[2] = t[2]; }
            if (nums(o->get("tolerance"), t, 1)) m->tolerance = t[0];
            if (nums(o->get("iterations"), t, 1)) m->iterations = (int)t[0];
            if (nums(o->get("impratio"), t, 1)) m->impratio = t[0];
            if (const std::string* s = o->get("integrator")) {
                if (*s == "RK4") m->integrator = ILQG_INT_RK4;
                else if (*s == "Euler") m->integrator = ILQG_INT_EULER;
                else return cx.fail("unsupported integrator " + *s);
            }
            if (const std::string* s = o->get("solver")) if (*s != "Newton") return cx.fail("only the Newton solver is supported");
            if (const std::string* s = o->get("cone")) if (*s != "pyramidal") return cx.fail("only the pyramidal cone is supported");
            if (m->impratio != 1) return cx.fail("impratio != 1 not supported");
            if (o->child("flag")) return cx.fail("<flag> not supported");  // o_solref etc. stay inert without the override flag
        }
        // world body
        m->nbody = 1;
        m->body_quat[0][0] = 1;
        bt.push_back({{0, 0, 0}, {1, 0, 0, 0}});
        const Elem* w = root->child("worldbody");
        if (!w) return cx.fail("no <worldbody>");
        for (auto& k : w->kids)
            if (k->name == "geom") { if (!addGeom(k.get(), 0)) return false; }
        for (auto& k : w->kids)
            if (k->name == "body") { if (!addBody(k.get(), 0)) return false; }
        if (root->child("equality") || root->child("tendon") || root->child("contact") || root->child("keyframe"))
            return cx.fail("equality/tendon/contact/keyframe sections not supported");
        if (const Elem* a = root->child("actuator")) {
            for (auto& k : a->kids) {
                if (k->name != "motor") return cx.fail("only <motor> actuators are supported");
                if (m->nu >= ILQG_MAXU) return cx.fail("too many actuators");
                int u = m->nu++;
                const std::string* jn = k->get("joint");
                if (!jn || !jntByName.count(*jn)) return cx.fail("motor without a known joint");
                int j = jntByName[*jn];
                if (m->jnt_type[j] == ILQG_JNT_FREE) return cx.fail("motor on a free joint");
                if (m->jnt_type[j] == ILQG_JNT_BALL) return cx.fail("motor on a ball joint is not supported");
                m->act_dofid[u] = m->jnt_dofadr[j];
                double t[6];
                m->act_gear[u] = nums(attr(k.get(), cx.defMotor, "gear"), t, 6) ? t[0] : 1.0;
                m->act_ctrllimited[u] = truthy(attr(k.get(), cx.defMotor, "ctrllimited"), false) ? 1 : 0;
                double r[2] = {0, 0};
                nums(attr(k.get(), cx.defMotor, "ctrlrange"), r, 2);
                m->act_ctrlrange[u][0] = r[0]; m->act_ctrlrange[u][1] = r[1];
            }
        }
        if (!inertiaFromGeoms()) return false;
        dofParents();
        if (!constantsAtQpos0()) return false;
        if (!makePairs()) return false;
        return true;
    }
};

}  // namespace

extern "C" int ilqg_compile_mjcf_string(const char* xml, ilqg_model* out, char* err, int errlen) {
    auto setErr = [&](const std::string& s) {
        if (err && errlen > 0) { snprintf(err, errlen, "%s", s.c_str()); }
    };
    if (!xml || !out) { setErr("null argument"); return ILQG_ERR_ARG; }
    std::string src(xml);
    XmlReader rd(src);
    auto root = rd.parse();
    if (!root) { setErr("XML: " + rd.err); return ILQG_ERR_MODEL; }
    Compiler c;
    c.m = out;
    if (!c.run(root.get())) { setErr("MJCF: " + c.cx.err); return ILQG_ERR_MODEL; }
    if (err && errlen > 0) err[0] = 0;
    return ILQG_OK;
}

extern "C" int ilqg_compile_mjcf(const char* path, ilqg_model* out, char* err, int errlen) {
    FILE* f = path ? fopen(path, "rb") : nullptr;
    if (!f) {
        if (err && errlen > 0) snprintf(err, errlen, "cannot open '%s'", path ? path : "(null)");
        return ILQG_ERR_IO;
    }
    std::string s;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return ilqg_compile_mjcf_string(s.c_str(), out, err, errlen);
}

extern "C" int ilqg_model_save(const char* path, const ilqg_model* m) {
    FILE* f = fopen(path, "wb");
    if (!f) return ILQG_ERR_IO;
    size_t n = fwrite(m, sizeof(*m), 1, f);
    fclose(f);
    return n == 1 ? ILQG_OK : ILQG_ERR_IO;
}

// Structural check of a table that did not come out of the compiler in this process (a file, a buffer handed over the ABI):
// counts within the fixed capacities, every index inside its range, tree order (parents before children), address
// consistency, and only the joint / geom / integrator kinds the kernels implement.  The kernels index with these values.
extern "C" int ilqg_model_validate(const ilqg_model* m, char* err, int errlen) {
    auto bad = [&](int code, const char* what, int idx) {
        if (err && errlen > 0) snprintf(err, errlen, "invalid model table: %s (index %d)", what, idx);
        return code;
    };
    if (!m) return bad(ILQG_ERR_ARG, "null table", 0);
    if (m->magic != ILQG_MODEL_MAGIC || m->version != ILQG_MODEL_VERSION) return bad(ILQG_ERR_MODEL, "bad magic / version", 0);
    if (m->nq < 1 || m->nq > ILQG_MAXQ || m->nv < 1 || m->nv > ILQG_MAXV || m->nu < 0 || m->nu > ILQG_MAXU || m->nbody < 2 ||
        m->nbody > ILQG_MAXBODY || m->njnt < 1 || m->njnt > ILQG_MAXJNT || m->ngeom < 0 || m->ngeom > ILQG_MAXGEOM || m->npair < 0 ||
        m->npair > ILQG_MAXPAIR)
        return bad(ILQG_ERR_MODEL, "a count is outside the table capacities", 0);
    if (!(m->timestep > 0) || m->iterations < 0 || m->ls_iterations < 0 || !(m->meaninertia > 0)) return bad(ILQG_ERR_MODEL, "option block", 0);
    if (m->integrator != ILQG_INT_EULER && m->integrator != ILQG_INT_RK4) return bad(ILQG_ERR_UNSUPPORTED, "integrator", m->integrator);
    int nq = 0, nv = 0;
    for (int j = 0; j < m->njnt; j++) {
        const int ty = m->jnt_type[j], b = m->jnt_bodyid[j];
        if (ty != ILQG_JNT_FREE && ty != ILQG_JNT_BALL && ty != ILQG_JNT_SLIDE && ty != ILQG_JNT_HINGE) return bad(ILQG_ERR_MODEL, "joint type", j);
        if (ty == ILQG_JNT_BALL && (m->jnt_limited[j] || m->jnt_stiffness[j] != 0 || m->body_jntnum[b] != 1))
            return bad(ILQG_ERR_UNSUPPORTED, "ball joint with a limit, a spring or sibling joints on its body", j);
        if (b < 1 || b >= m->nbody) return bad(ILQG_ERR_MODEL, "jnt_bodyid", j);
        if (m->jnt_qposadr[j] != nq || m->jnt_dofadr[j] != nv) return bad(ILQG_ERR_MODEL, "jnt_qposadr / jnt_dofadr not cumulative", j);
        nq += ty == ILQG_JNT_FREE ? 7 : (ty == ILQG_JNT_BALL ? 4 : 1);
        nv += ty == ILQG_JNT_FREE ? 6 : (ty == ILQG_JNT_BALL ? 3 : 1);
    }
    if (nq != m->nq || nv != m->nv) return bad(ILQG_ERR_MODEL, "nq / nv do not match the joints", 0);
    if (m->body_parentid[0] != 0) return bad(ILQG_ERR_MODEL, "body 0 must be the world", 0);
    for (int b = 1; b < m->nbody; b++) {
        if (m->body_parentid[b] < 0 || m->body_parentid[b] >= b) return bad(ILQG_ERR_MODEL, "body_parentid (parents precede children)", b);
        if (m->body_rootid[b] < 1 || m->body_rootid[b] > b) return bad(ILQG_ERR_MODEL, "body_rootid", b);
        if (m->body_jntnum[b] < 0 || m->body_dofnum[b] < 0) return bad(ILQG_ERR_MODEL, "body_jntnum / body_dofnum", b);
        if (m->body_jntnum[b] > 0 && (m->body_jntadr[b] < 0 || m->body_jntadr[b] + m->body_jntnum[b] > m->njnt)) return bad(ILQG_ERR_MODEL, "body_jntadr", b);
        if (m->body_dofnum[b] > 0 && (m->body_dofadr[b] < 0 || m->body_dofadr[b] + m->body_dofnum[b] > m->nv)) return bad(ILQG_ERR_MODEL, "body_dofadr", b);
        if (!(m->body_mass[b] >= 0)) return bad(ILQG_ERR_MODEL, "body_mass", b);
    }
    for (int i = 0; i < m->nv; i++) {
        if (m->dof_bodyid[i] < 1 || m->dof_bodyid[i] >= m->nbody) return bad(ILQG_ERR_MODEL, "dof_bodyid", i);
        if (m->dof_jntid[i] < 0 || m->dof_jntid[i] >= m->njnt) return bad(ILQG_ERR_MODEL, "dof_jntid", i);
        if (m->dof_parentid[i] < -1 || m->dof_parentid[i] >= i) return bad(ILQG_ERR_MODEL, "dof_parentid", i);
    }
    for (int g = 0; g < m->ngeom; g++) {
        const int ty = m->geom_type[g];
        if (ty != ILQG_GEOM_PLANE && ty != ILQG_GEOM_SPHERE && ty != ILQG_GEOM_CAPSULE) return bad(ILQG_ERR_UNSUPPORTED, "geom type", g);
        if (m->geom_bodyid[g] < 0 || m->geom_bodyid[g] >= m->nbody) return bad(ILQG_ERR_MODEL, "geom_bodyid", g);
    }
    for (int p = 0; p < m->npair; p++) {
        if (m->pair_geom1[p] < 0 || m->pair_geom1[p] >= m->ngeom || m->pair_geom2[p] < 0 || m->pair_geom2[p] >= m->ngeom)
            return bad(ILQG_ERR_MODEL, "pair_geom", p);
        if (m->pair_condim[p] != 1 && m->pair_condim[p] != 3) return bad(ILQG_ERR_UNSUPPORTED, "pair_condim (1 and 3 are implemented)", p);
        if (m->geom_type[m->pair_geom2[p]] == ILQG_GEOM_PLANE) return bad(ILQG_ERR_MODEL, "a plane must be geom1 of its pair", p);
    }
    for (int a = 0; a < m->nu; a++)
        if (m->act_dofid[a] < 0 || m->act_dofid[a] >= m->nv) return bad(ILQG_ERR_MODEL, "act_dofid", a);
    if (err && errlen > 0) err[0] = 0;
    return ILQG_OK;
}

extern "C" int ilqg_model_load(const char* path, ilqg_model* m) {
    if (!path || !m) return ILQG_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return ILQG_ERR_IO;
    size_t n = fread(m, sizeof(*m), 1, f);
    const bool more = n == 1 && fgetc(f) != EOF;   // a file of another size is not one of our tables
    fclose(f);
    if (n != 1 || more) return ILQG_ERR_IO;
    return ilqg_model_validate(m, nullptr, 0);
}

extern "C" int ilqg_model_sizeof(void) { return (int)sizeof(ilqg_model); }

// byte offset and element count of a named field of ilqg_model (for bindings that treat the table as bytes)
extern "C" int ilqg_model_field(const char* name, int* offset, int* count, int* is_double) {
#define F(field, n, dbl)                                   \
    if (!strcmp(name, #field)) {                           \
        *offset = (int)offsetof(ilqg_model, field);        \
        *count = (n);                                      \
        *is_double = (dbl);                                \
        return ILQG_OK;                                    \
    }
    if (!name || !offset || !count || !is_double) return ILQG_ERR_ARG;
    F(timestep, 1, 1) F(gravity, 3, 1) F(tolerance, 1, 1) F(meaninertia, 1, 1) F(integrator, 1, 0) F(iterations, 1, 0)
    F(npair, 1, 0) F(body_parentid, ILQG_MAXBODY, 0) F(body_mass, ILQG_MAXBODY, 1) F(body_ipos, ILQG_MAXBODY * 3, 1)
    F(body_inertia, ILQG_MAXBODY * 6, 1) F(body_invweight0, ILQG_MAXBODY * 2, 1) F(body_pos, ILQG_MAXBODY * 3, 1)
    F(jnt_type, ILQG_MAXJNT, 0) F(jnt_limited, ILQG_MAXJNT, 0) F(jnt_range, ILQG_MAXJNT * 2, 1) F(jnt_stiffness, ILQG_MAXJNT, 1)
    F(jnt_axis, ILQG_MAXJNT * 3, 1) F(jnt_pos, ILQG_MAXJNT * 3, 1) F(jnt_solimp, ILQG_MAXJNT * 5, 1) F(qpos0, ILQG_MAXQ, 1)
    F(dof_armature, ILQG_MAXV, 1) F(dof_damping, ILQG_MAXV, 1) F(dof_invweight0, ILQG_MAXV, 1) F(dof_parentid, ILQG_MAXV, 0)
    F(geom_type, ILQG_MAXGEOM, 0) F(geom_size, ILQG_MAXGEOM * 3, 1) F(geom_pos, ILQG_MAXGEOM * 3, 1) F(geom_quat, ILQG_MAXGEOM * 4, 1)
    F(pair_geom1, ILQG_MAXPAIR, 0) F(pair_geom2, ILQG_MAXPAIR, 0) F(pair_condim, ILQG_MAXPAIR, 0) F(pair_margin, ILQG_MAXPAIR, 1)
    F(pair_friction, ILQG_MAXPAIR, 1) F(pair_solref, ILQG_MAXPAIR * 2, 1) F(pair_solimp, ILQG_MAXPAIR * 5, 1)
    F(act_dofid, ILQG_MAXU, 0) F(act_gear, ILQG_MAXU, 1) F(act_ctrllimited, ILQG_MAXU, 0) F(act_ctrlrange, ILQG_MAXU * 2, 1)
#undef F
    return ILQG_ERR_ARG;
}
