#!/usr/bin/env python3
"""Generate robot-specific code from ompl_b200/robots/<robot>.json (no reference access needed).

Emits
  ompl_b200/csrc/robots/<robot>_gen.cuh   link-by-link forward kinematics for the CUDA engine:
        one `frame<K>()` per actuated joint (parent frame x constant origin x joint motion, with
        zero / unit coefficients folded away at generation time) plus sphere / link / pair tables.
  oracle/robots_gen.h                      plain C tables for the table-driven CPU oracle.

This is our own structured generator (frames -> spheres), not the reference's flat traced program
(external/vamp/src/impl/vamp/robots/panda.hh:5226-5539 is the behaviour it must reproduce).
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

ROBOT_IDS = {"panda": 0, "fetch": 1, "ur5": 2, "baxter": 3}
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, ".."))
EPS = 1e-9


def f32(v: float) -> str:
    """C float literal of the float32 nearest to the double v (single cast, like the reference's
    scalar-broadcast operators)."""
    x = float(np.float32(v))
    s = repr(x)
    if "e" not in s and "." not in s and "inf" not in s:
        s += ".0"
    return s + "f"


class Emitter:
    """Tiny symbolic layer: values are python floats (constants) or C expression strings."""

    def __init__(self):
        self.lines = []
        self.n = 0

    @staticmethod
    def is_c(v):
        return isinstance(v, float)

    def mul(self, a, b):
        if self.is_c(a) and self.is_c(b):
            return a * b
        if self.is_c(b):
            a, b = b, a
        if self.is_c(a):
            if abs(a) < EPS:
                return 0.0
            if abs(a - 1) < EPS:
                return b
            if abs(a + 1) < EPS:
                return self.neg(b)
            return f"({f32(a)} * {b})"
        return f"({a} * {b})"

    def neg(self, a):
        if self.is_c(a):
            return -a
        if a.startswith("(-") and a.endswith(")") and a.count("(") == 1:
            return a[2:-1]
        return f"(-{a})"

    def add(self, a, b):
        if self.is_c(a) and self.is_c(b):
            return a + b
        if self.is_c(a):
            a, b = b, a
        if self.is_c(b):
            if abs(b) < EPS:
                return a
            return f"({a} + {f32(b)})"
        return f"({a} + {b})"

    def sum(self, terms):
        acc = 0.0
        for t in terms:
            acc = self.add(acc, t)
        return acc

    def lit(self, v):
        return f32(v) if self.is_c(v) else v

    def assign(self, name, v):
        self.lines.append(f"{name} = {self.lit(v)};")


def gen_frame(em: Emitter, fr: dict, parent_is_world: bool):
    """Emit code computing frame F (R row-major r[9], t[3]) from parent P."""
    rc = np.asarray(fr["R"], dtype=np.float64)
    tc = np.asarray(fr["t"], dtype=np.float64)
    if parent_is_world:
        pr = [[float(i == j) for j in range(3)] for i in range(3)]
        pt = [0.0, 0.0, 0.0]
    else:
        pr = [[f"P.r[{3 * i + j}]" for j in range(3)] for i in range(3)]
        pt = [f"P.t[{i}]" for i in range(3)]
    # A = Pr * Rc ; t = pt + Pr * tc
    a = [[em.sum([em.mul(pr[i][k], float(rc[k][j])) for k in range(3)]) for j in range(3)] for i in range(3)]
    t = [em.add(em.sum([em.mul(pr[i][k], float(tc[k])) for k in range(3)]), pt[i]) for i in range(3)]
    axis = np.asarray(fr["axis"], dtype=np.float64)
    ai = int(np.argmax(np.abs(axis)))
    sgn = float(np.sign(axis[ai]))
    assert abs(abs(axis[ai]) - 1) < 1e-12 and np.abs(axis).sum() - 1 < 1e-12, "axis-aligned joints only"
    if fr["type"] in ("revolute", "continuous"):
        # materialise A's two rotating columns as temporaries when they are expressions
        s = "s" if sgn > 0 else "(-s)"
        u, v = [(1, 2), (2, 0), (0, 1)][ai]  # rotation about axis ai mixes columns u, v: u' = c u + s v ; v' = -s u + c v
        r = [[None] * 3 for _ in range(3)]
        for i in range(3):
            cu, cv = a[i][u], a[i][v]
            r[i][ai] = a[i][ai]
            r[i][u] = em.add(em.mul(cu, "c"), em.mul(cv, s))
            r[i][v] = em.add(em.mul(em.neg(cu), s), em.mul(cv, "c"))
    elif fr["type"] == "prismatic":
        r = a
        t = [em.add(t[i], em.mul(a[i][ai], "q" if sgn > 0 else "(-q)")) for i in range(3)]
    else:
        raise ValueError(fr["type"])
    for i in range(3):
        for j in range(3):
            em.assign(f"F.r[{3 * i + j}]", r[i][j])
    for i in range(3):
        em.assign(f"F.t[{i}]", t[i])


def model_fk(model: dict, q):
    """world pose (R, t) of every frame of the JSON model for one configuration (float64; generator-side estimates only)."""
    frames = model["frames"]
    T = [None] * len(frames)
    T[0] = (np.eye(3), np.zeros(3))
    for fr in frames[1:]:
        pr, pt = T[fr["parent"]]
        a = pr @ np.asarray(fr["R"], dtype=np.float64)
        t = pt + pr @ np.asarray(fr["t"], dtype=np.float64)
        ax = np.asarray(fr["axis"], dtype=np.float64)
        v = q[fr["q_index"]]
        if fr["type"] in ("revolute", "continuous"):
            u = ax / np.linalg.norm(ax)
            k = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
            a = a @ (np.eye(3) + math.sin(v) * k + (1 - math.cos(v)) * (k @ k))
        else:
            t = t + a @ (ax * v)
        T[fr["id"]] = (a, t)
    return T


def pair_groups(model: dict, link_rows: list, pairs: list):
    """Self-collision pairs that share one link x and whose other links live on ONE frame can be ruled out together: a sphere around
    one member's bounding centre that contains every member's bounding sphere (same frame: a rigid set) stands for all of them. A
    group is worth a test of its own when it usually fails to touch (estimated on random configurations) -- x's sphere not touching
    the group sphere implies it touches no member, so skipping the members then changes nothing.
    -> list of dict(x, ref, r, members=[pair index ...])"""
    lo = np.float32(model["s_a"]).astype(np.float64)
    hi = lo + np.float32(model["s_m"]).astype(np.float64)
    rng = np.random.default_rng(20260202)
    qs = rng.uniform(lo, hi, (300, len(lo)))
    centre = np.zeros((len(qs), len(link_rows), 3))
    for i, q in enumerate(qs):
        T = model_fk(model, q)
        for l, row in enumerate(link_rows):
            r_, t_ = T[row["frame"]]
            centre[i, l] = r_ @ np.asarray(row["coarse"]["c"]) + t_
    rad = np.array([row["coarse"]["r"] for row in link_rows])
    cand = {}
    for pi, (a, b) in enumerate(pairs):
        cand.setdefault((a, link_rows[b]["frame"]), []).append((pi, b))
        cand.setdefault((b, link_rows[a]["frame"]), []).append((pi, a))
    scored = []
    for (x, f), mem in cand.items():
        if len(mem) < 3:
            continue
        best = None
        for _, ref in mem:
            cr = np.asarray(link_rows[ref]["coarse"]["c"])
            rg = max(np.linalg.norm(np.asarray(link_rows[m]["coarse"]["c"]) - cr) + rad[m] for _, m in mem)
            if best is None or rg < best[0]:
                best = (rg, ref)
        rg, ref = best[0] + 1e-5, best[1]   # (margin: far above the rounding of the device-side test)
        p_touch = float((np.linalg.norm(centre[:, x] - centre[:, ref], axis=1) < rad[x] + rg).mean())
        saved = (1 - p_touch) * len(mem) - 1
        if saved > 0.75:
            scored.append((saved, x, ref, rg, [pi for pi, _ in mem]))
    scored.sort(key=lambda g: -g[0])
    used, groups = set(), []
    for saved, x, ref, rg, mem in scored:
        mem = [pi for pi in mem if pi not in used]
        if len(mem) < 3:
            continue
        used.update(mem)
        groups.append(dict(x=x, ref=ref, r=float(rg), members=mem))
    return groups


def gen_cuda(model: dict) -> str:
    name = model["name"]
    cname = name.capitalize()
    frames = model["frames"]
    links = model["links"]
    # process links in frame order (base -> tip) so frames stream; keep the reference's env order as data
    order = sorted(range(len(links)), key=lambda i: (links[i]["frame"], -i))
    fine_sorted = []
    link_rows = []
    for li in order:
        l = links[li]
        b = len(fine_sorted)
        for f in l["fine"]:
            fine_sorted.append((f["c"], f["r"], f["id"]))
        link_rows.append(dict(ref_index=li, name=l["name"], frame=l["frame"], begin=b, end=len(fine_sorted),
                              coarse=l["coarse"]))
    new_index = {row["ref_index"]: i for i, row in enumerate(link_rows)}
    # self pairs as (link on the earlier frame, link on the later frame), grouped by the later frame so that
    # stage 2 can stream frames base -> tip and only keep the few "earlier" frames in saved slots
    pairs = []
    for a, b in model["self_pairs"]:
        a, b = new_index[a], new_index[b]
        if link_rows[a]["frame"] > link_rows[b]["frame"]:
            a, b = b, a
        pairs.append((a, b))
    pairs.sort(key=lambda p: (link_rows[p[1]]["frame"], link_rows[p[0]]["frame"], p))
    nfr = len(frames)
    # frames form chains hanging off the world frame (one per arm); frame_branch[k] = first frame of k's chain
    frame_branch = [0] * nfr
    for fr in frames[1:]:
        assert fr["parent"] in (0, fr["id"] - 1), "every frame must continue the previous chain or start a new one at the world frame"
        frame_branch[fr["id"]] = fr["id"] if fr["parent"] == 0 else frame_branch[fr["parent"]]
    frame_link_begin = [next((i for i, r in enumerate(link_rows) if r["frame"] >= k), len(link_rows)) for k in range(nfr)] + [len(link_rows)]
    pair_frame_begin = [next((i for i, p in enumerate(pairs) if link_rows[p[1]]["frame"] >= k), len(pairs)) for k in range(nfr)] + [len(pairs)]
    a_frames = sorted({link_rows[a]["frame"] for a, _ in pairs} - {0})
    save_slot = [a_frames.index(k) if k in a_frames else -1 for k in range(nfr)]
    # which saved slot holds the earlier link's frame: -2 = world, -1 = same frame as the later link
    pair_slot = []
    for a, b in pairs:
        fa, fb = link_rows[a]["frame"], link_rows[b]["frame"]
        pair_slot.append(-2 if fa == 0 else (-1 if fa == fb else a_frames.index(fa)))

    lo = [float(np.float32(a)) for a in model["s_a"]]
    hi = [float(np.float32(np.float32(a) + np.float32(m))) for a, m in zip(model["s_a"], model["s_m"])]
    # The state check as a task list, so that ONE loop (one copy of the frame code) runs all of it: the forward pass over the frames
    # (bounding-sphere centres), the links against the environment tip -> base (the reference's order), the self-collision pairs
    # (first the frame of the link whose fine spheres run in the outer loop -- the one with fewer spheres --, then the other's).
    T_FK, T_ENV, T_SELF_A, T_SELF_B, T_GROUP = 0, 1, 2, 3, 4
    tasks = [(T_FK, k, frame_link_begin[k], frame_link_begin[k + 1]) for k in range(nfr)]
    tasks += [(T_ENV, link_rows[l]["frame"], l, 0) for l in range(len(link_rows) - 1, -1, -1)]

    def pair_tasks(a, b):
        na, nb = link_rows[a]["end"] - link_rows[a]["begin"], link_rows[b]["end"] - link_rows[b]["begin"]
        o, i = (a, b) if na <= nb else (b, a)
        return [(T_SELF_A, link_rows[o]["frame"], o, i), (T_SELF_B, link_rows[i]["frame"], o, i)]

    groups = pair_groups(model, link_rows, pairs)
    grouped = {pi for g in groups for pi in g["members"]}
    for pi, (a, b) in enumerate(pairs):
        if pi not in grouped:
            tasks += pair_tasks(a, b)
    for gi, g in enumerate(groups):  # a group test, then its pairs (skipped together when the group sphere is not touched)
        tasks.append((T_GROUP, 0, g["x"], gi))
        for pi in g["members"]:
            tasks += pair_tasks(*pairs[pi])
    att = model["attachment"]
    att_links = sorted((new_index[a] for a in att["links"]), key=lambda i: (link_rows[i]["frame"], i))

    out = []
    w = out.append
    w(f"// GENERATED by tools/gen_robot_code.py from ompl_b200/robots/{name}.json -- do not edit.")
    w(f"// Link-by-link forward kinematics + collision-sphere tables for the '{name}' robot.")
    w(f"// Behaviour to reproduce: vamp::robots::{cname}::fkcc (external/vamp/src/impl/vamp/robots/{name}.hh).")
    w("#pragma once")
    w('#include "../robot_common.cuh"')
    w("")
    w("namespace b200mp {")
    w("")
    nl, nf, npairs = len(link_rows), len(fine_sorted), len(pairs)
    w(f"struct {cname}Tables {{")
    w(f"    float4 coarse[{nl}];        // frame-local centre + radius of each link's bounding sphere")
    w(f"    float4 fine[{nf}];          // frame-local centre + radius, grouped by link")
    w(f"    short link_begin[{nl + 1}]; // fine-sphere range of link l = [link_begin[l], link_begin[l+1])")
    w(f"    short pair_a[{npairs}], pair_b[{npairs}], pair_slot[{npairs}];  // earlier-frame link, later-frame link, saved slot of the earlier frame (-2 world, -1 same)")
    w(f"    short frame_link_begin[{nfr + 1}];  // links living on frame k = [frame_link_begin[k], frame_link_begin[k+1])")
    w(f"    short pair_frame_begin[{nfr + 1}];  // pairs whose later link lives on frame k")
    w(f"    short save_slot[{nfr}];             // slot frame k is saved to for later pairs, or -1")
    w(f"    short link_frame[{nl}];             // frame link l lives on")
    w(f"    short frame_branch[{nfr}];          // first frame of the chain (arm) frame k belongs to; its parent is the world frame")
    w(f"    short att_links[{len(att_links)}];  // links tested against an attachment (Robot::fkcc_attach), in frame order")
    w(f"    RobotTask tasks[{len(tasks)}];       // the state check as one list of steps (see RobotTask in robot_common.cuh)")
    w(f"    short group_ref[{max(1, len(groups))}], group_skip[{max(1, len(groups))}];  // pair groups: member whose bounding centre carries the group sphere; tasks to skip")
    w(f"    float group_r[{max(1, len(groups))}];              // radius of the group sphere (contains every member's bounding sphere)")
    w("};")
    w("")
    w(f"__constant__ {cname}Tables k{cname}Tables = {{")
    w("    {" + ", ".join("{%s, %s, %s, %s}" % (f32(r["coarse"]["c"][0]), f32(r["coarse"]["c"][1]),
                                                 f32(r["coarse"]["c"][2]), f32(r["coarse"]["r"])) for r in link_rows) + "},")
    w("    {" + ",\n     ".join("{%s, %s, %s, %s}" % (f32(c[0]), f32(c[1]), f32(c[2]), f32(r)) for c, r, _ in fine_sorted) + "},")
    w("    {" + ", ".join(str(r["begin"]) for r in link_rows) + f", {nf}" + "},")
    w("    {" + ", ".join(str(a) for a, _ in pairs) + "},")
    w("    {" + ", ".join(str(b) for _, b in pairs) + "},")
    w("    {" + ", ".join(str(v) for v in pair_slot) + "},")
    w("    {" + ", ".join(str(v) for v in frame_link_begin) + "},")
    w("    {" + ", ".join(str(v) for v in pair_frame_begin) + "},")
    w("    {" + ", ".join(str(v) for v in save_slot) + "},")
    w("    {" + ", ".join(str(r["frame"]) for r in link_rows) + "},")
    w("    {" + ", ".join(str(v) for v in frame_branch) + "},")
    w("    {" + ", ".join(str(v) for v in att_links) + "},")
    w("    {" + ", ".join("{%d, %d, %d, %d}" % t for t in tasks) + "},")
    w("    {" + (", ".join(str(g["ref"]) for g in groups) or "0") + "},")
    w("    {" + (", ".join(str(2 * len(g["members"])) for g in groups) or "0") + "},")
    w("    {" + (", ".join(f32(g["r"]) for g in groups) or "0.0f") + "},")
    w("};")
    w("")
    w(f"struct {cname} {{")
    w(f"    static constexpr int kDim = {model['dimension']};")
    w(f"    static constexpr int kFrames = {len(frames)};   // frame 0 = world; frames form a chain base -> tip")
    w(f"    static constexpr int kLinks = {nl};")
    w(f"    static constexpr int kFine = {nf};")
    w(f"    static constexpr int kPairs = {npairs};")
    w(f"    static constexpr int kResolution = {model['resolution']};")
    w(f"    static constexpr int kSlots = {max(1, len(a_frames))};   // saved frames needed by the self-collision stage")
    w(f"    static constexpr int kRobotId = {ROBOT_IDS[name]};")
    w(f"    static constexpr int kTasks = {len(tasks)};")
    w(f"    static constexpr int kAttLinks = {len(att_links)};   // links an attachment is tested against")
    w(f"    static constexpr int kEeFrame = {att['ee_frame']};   // moving frame the end effector ({cname}::eefk) is rigidly attached to")
    w("    static constexpr float kLo[kDim] = {" + ", ".join(f32(v) for v in lo) + "};")
    w("    static constexpr float kHi[kDim] = {" + ", ".join(f32(v) for v in hi) + "};")
    w(f"    using Tables = {cname}Tables;")
    w(f"    __device__ static __forceinline__ const Tables &tables() {{ return k{cname}Tables; }}")
    w("")
    w("    // F = P * origin(K) * motion(q).  s, c = vamp_sin(q), vamp_cos(q) for revolute joints; q raw for prismatic.")
    w("    template <int K>")
    w("    __device__ static __forceinline__ void frame(const Frame &P, float q, float s, float c, Frame &F);")
    w("")
    w("    // sin / cos of every revolute joint (the reference's polynomials), computed once per state")
    w("    __device__ static __forceinline__ void trig(const float (&q)[kDim], float (&sn)[kDim], float (&cs)[kDim]) {")
    w("#pragma unroll")
    w("        for (int j = 0; j < kDim; ++j) {")
    w("            if (joint_revolute(j)) { sn[j] = vamp_sin(q[j]); cs[j] = vamp_cos(q[j]); } else { sn[j] = 0.0f; cs[j] = 1.0f; }")
    w("        }")
    w("    }")
    w("    __host__ __device__ static constexpr bool joint_revolute(int j) {")
    w("        constexpr bool v[kDim] = {" + ", ".join("true" if fr["type"] in ("revolute", "continuous") else "false" for fr in sorted(frames[1:], key=lambda f: f["q_index"])) + "};")
    w("        return v[j];")
    w("    }")
    w("    // one base -> tip pass: st(l, x, y, z) receives the world-space centre of every link's bounding sphere, tip the last frame")
    w("    // (same arithmetic as frame_apply on the streamed frames; zero coefficients dropped, which leaves the values unchanged)")
    w("    template <class St>")
    w("    __device__ static __forceinline__ void fk_coarse(const float (&q)[kDim], const float (&sn)[kDim], const float (&cs)[kDim], Frame &tip, St st);")
    w("    // E = P * (constant end-effector offset), P = frame kEeFrame: the pose Robot::eefk returns, in which attachments are posed")
    w("    __device__ static __forceinline__ void ee_pose(const Frame &P, Frame &E);")
    w("    // step the streaming frame from k-1 to k (k >= 1); uniform switch, static joint indices")
    w("    __device__ static __forceinline__ void advance(int k, Frame &F, const float (&q)[kDim], const float (&sn)[kDim], const float (&cs)[kDim]);")
    w("    // The same step OUT OF LINE, frame in and out by value (registers): the one copy of the frame code every caller shares, so that")
    w("    // the kernels' hot code fits the instruction cache. s, c = the joint's sin / cos; a prismatic joint passes its value q as s.")
    w("    // (a template only for its linkage: the header is included by several translation units)")
    w("    template <int = 0>")
    w("    __device__ static __noinline__ Frame step(int k, Frame P, float s, float c);")
    w("    // ... and inline, for the one place a kernel's task loop steps frames (validate.cuh: check_state_grid)")
    w("    __device__ static __forceinline__ Frame step_inline(int k, const Frame &P, float s, float c);")
    w("};")
    w("")
    for fr in frames[1:]:
        em = Emitter()
        gen_frame(em, fr, parent_is_world=(fr["parent"] == 0))
        w(f"// frame {fr['id']}: joint '{fr['joint']}' ({fr['type']}), parent frame {fr['parent']}")
        w("template <>")
        w(f"__device__ __forceinline__ void {cname}::frame<{fr['id']}>(const Frame &P, float q, float s, float c, Frame &F) {{")
        w("    (void)P; (void)q; (void)s; (void)c;")
        for ln in em.lines:
            w("    " + ln)
        w("}")
        w("")
    em = Emitter()
    er = np.asarray(att["R"], dtype=np.float64)
    et = np.asarray(att["t"], dtype=np.float64)
    pr = [[f"P.r[{3 * i + j}]" for j in range(3)] for i in range(3)]
    w(f"__device__ __forceinline__ void {cname}::ee_pose(const Frame &P, Frame &E) {{")
    for i in range(3):
        for j in range(3):
            w(f"    E.r[{3 * i + j}] = {em.lit(em.sum([em.mul(pr[i][k], float(er[k][j])) for k in range(3)]))};")
    for i in range(3):
        w(f"    E.t[{i}] = {em.lit(em.add(em.sum([em.mul(pr[i][k], float(et[k])) for k in range(3)]), f'P.t[{i}]'))};")
    w("}")
    w("")

    def apply_expr(fname, row, c):
        acc = f"{fname}.t[{row}]"
        for j in (2, 1, 0):
            cj = float(np.float32(c[j]))
            if cj != 0.0:
                acc = f"fmaf({fname}.r[{3 * row + j}], {f32(cj)}, {acc})"
        return acc

    w("template <class St>")
    w(f"__device__ __forceinline__ void {cname}::fk_coarse(const float (&q)[kDim], const float (&sn)[kDim], const float (&cs)[kDim], Frame &tip, St st) {{")
    w("    (void)q; (void)sn; (void)cs;")
    w("    Frame F0;")
    w("    frame_identity(F0);")
    for fr in frames:
        k = fr["id"]
        if k > 0:
            qi = fr["q_index"]
            w(f"    Frame F{k};")
            w(f"    frame<{k}>(F{fr['parent']}, q[{qi}], sn[{qi}], cs[{qi}], F{k});")
        for li, r in enumerate(link_rows):
            if r["frame"] != k:
                continue
            c = r["coarse"]["c"]
            if k == 0:
                w(f"    st({li}, {f32(c[0])}, {f32(c[1])}, {f32(c[2])});")
            else:
                w(f"    st({li}, {apply_expr(f'F{k}', 0, c)}, {apply_expr(f'F{k}', 1, c)}, {apply_expr(f'F{k}', 2, c)});")
    w(f"    tip = F{len(frames) - 1};")
    w("}")
    w("")
    w(f"__device__ __forceinline__ void {cname}::advance(int k, Frame &F, const float (&q)[kDim], const float (&sn)[kDim], const float (&cs)[kDim]) {{")
    w("    const Frame P = F;")
    w("    switch (k) {")
    for fr in frames[1:]:
        qi = fr["q_index"]
        w(f"        case {fr['id']}: frame<{fr['id']}>(P, q[{qi}], sn[{qi}], cs[{qi}], F); break;")
    w("        default: break;")
    w("    }")
    w("}")
    w("")
    w(f"__device__ __forceinline__ Frame {cname}::step_inline(int k, const Frame &P, float s, float c) {{")
    w("    Frame F = P;")
    w("    switch (k) {")
    for fr in frames[1:]:
        w(f"        case {fr['id']}: frame<{fr['id']}>(P, s, s, c, F); break;")
    w("        default: break;")
    w("    }")
    w("    return F;")
    w("}")
    w("")
    w("template <int>")
    w(f"__device__ __noinline__ Frame {cname}::step(int k, Frame P, float s, float c) {{ return step_inline(k, P, s, c); }}")
    w("")
    w("}  // namespace b200mp")
    return "\n".join(out) + "\n", link_rows, fine_sorted, pairs


def gen_oracle_tables(models: list) -> str:
    out = []
    w = out.append
    w("/* GENERATED by tools/gen_robot_code.py from ompl_b200/robots/*.json -- do not edit.")
    w(" * TEST INFRASTRUCTURE ONLY: robot tables for the table-driven CPU oracle (oracle/oracle_validate.cpp).")
    w(" * Links are listed in the reference's own environment-test order; doubles keep full URDF precision. */")
    w("#pragma once")
    w("typedef struct { int parent; int q_index; int type; /* 0 revolute, 1 prismatic */ double axis[3]; double R[9]; double t[3]; } OrFrame;")
    w("typedef struct { double c[3]; double r; } OrSphere;")
    w("typedef struct { int frame; OrSphere coarse; int fine_begin, fine_end; } OrLink;")
    w("typedef struct { const char *name; int dim, n_frames, n_links, n_fine, n_pairs, resolution;")
    w("                 const OrFrame *frames; const OrLink *links; const OrSphere *fine; const int *fine_id; const int (*pairs)[2];")
    w("                 const float *lo; const float *hi;")
    w("                 int ee_frame; const double *ee_R; const double *ee_t; int n_att_links; const int *att_links; } OrRobot;")
    for m in models:
        n = m["name"]
        frames = m["frames"]
        w(f"static const OrFrame {n}_frames[] = {{")
        for fr in frames:
            if fr["id"] == 0:
                w("  {-1, -1, 0, {0,0,1}, {1,0,0,0,1,0,0,0,1}, {0,0,0}},")
                continue
            r = np.asarray(fr["R"]).reshape(-1)
            w("  {%d, %d, %d, {%s}, {%s}, {%s}}," % (
                fr["parent"], fr["q_index"], 0 if fr["type"] in ("revolute", "continuous") else 1,
                ", ".join(repr(float(v)) for v in fr["axis"]),
                ", ".join(repr(float(v)) for v in r), ", ".join(repr(float(v)) for v in fr["t"])))
        w("};")
        fine = []
        w(f"static const OrLink {n}_links[] = {{")
        for l in m["links"]:
            b = len(fine)
            fine.extend(l["fine"])
            c = l["coarse"]
            w("  {%d, {{%s}, %s}, %d, %d}," % (l["frame"], ", ".join(repr(float(v)) for v in c["c"]), repr(float(c["r"])), b, len(fine)))
        w("};")
        w(f"static const OrSphere {n}_fine[] = {{")
        for f in fine:
            w("  {{%s}, %s}," % (", ".join(repr(float(v)) for v in f["c"]), repr(float(f["r"]))))
        w("};")
        w(f"static const int {n}_fine_id[] = {{" + ", ".join(str(f["id"]) for f in fine) + "};  /* reference sphere index */")
        w(f"static const int {n}_pairs[][2] = {{" + ", ".join("{%d,%d}" % (a, b) for a, b in m["self_pairs"]) + "};")
        lo = [float(np.float32(a)) for a in m["s_a"]]
        hi = [float(np.float32(np.float32(a) + np.float32(mm))) for a, mm in zip(m["s_a"], m["s_m"])]
        w(f"static const float {n}_lo[] = {{" + ", ".join(f32(v) for v in lo) + "};")
        w(f"static const float {n}_hi[] = {{" + ", ".join(f32(v) for v in hi) + "};")
        at = m["attachment"]
        w(f"static const double {n}_ee_R[] = {{" + ", ".join(repr(float(v)) for v in np.asarray(at["R"]).reshape(-1)) + "};")
        w(f"static const double {n}_ee_t[] = {{" + ", ".join(repr(float(v)) for v in at["t"]) + "};")
        w(f"static const int {n}_att_links[] = {{" + ", ".join(str(v) for v in at["links"]) + "};  /* reference order */")
        w(f"static const OrRobot {n}_robot = {{\"{n}\", {m['dimension']}, {len(frames)}, {len(m['links'])}, {len(fine)}, "
          f"{len(m['self_pairs'])}, {m['resolution']}, {n}_frames, {n}_links, {n}_fine, {n}_fine_id, {n}_pairs, {n}_lo, {n}_hi, "
          f"{at['ee_frame']}, {n}_ee_R, {n}_ee_t, {len(at['links'])}, {n}_att_links}};")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    models = []
    for r in ROBOT_IDS:
        p = os.path.join(ROOT, "ompl_b200", "robots", f"{r}.json")
        if not os.path.exists(p):
            continue
        m = json.load(open(p))
        models.append(m)
        code, *_ = gen_cuda(m)
        os.makedirs(os.path.join(ROOT, "ompl_b200", "csrc", "robots"), exist_ok=True)
        with open(os.path.join(ROOT, "ompl_b200", "csrc", "robots", f"{r}_gen.cuh"), "w") as f:
            f.write(code)
    with open(os.path.join(ROOT, "oracle", "robots_gen.h"), "w") as f:
        f.write(gen_oracle_tables(models))
    print("generated", [m["name"] for m in models])
