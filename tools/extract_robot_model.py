#!/usr/bin/env python3
"""Dev-time tool (runs only where /root/reference exists): derive a structured robot model.

Reads, as DATA,
  * the spherized URDF        external/vamp/resources/<robot>/<robot>_spherized.urdf
  * the reference's traced FK  external/vamp/src/impl/vamp/robots/<robot>.hh  (Robot::fkcc)
and writes ompl_b200/robots/<robot>.json: kinematic frames (one per actuated joint, constant
origin transform + joint motion), per-link fine spheres in frame-local coordinates, the per-link
bounding ("coarse") spheres, the order links are tested against the environment, and the
allowed self-collision link pairs.

The URDF gives frames and fine spheres. Two things exist only inside the reference's generated
code and are recovered from it numerically/structurally:
  * the coarse spheres: the trace is evaluated in float64 at random configurations and each
    coarse centre is solved for in its link frame (must be constant to 1e-9);
  * the link -> {coarse, fine} grouping, env order and self-collision pairs: parsed from the
    `sphere_environment_in_collision` / `sphere_sphere_self_collision` call structure
    (robots/panda.hh:5629-10260).
The URDF-derived FK is then checked against the evaluated trace for every fine sphere (< 1e-6 m).
No reference code is emitted; tools/gen_robot_code.py turns the JSON into our own link-by-link FK.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("OMPL_REF", "/root/reference")


# ----------------------------------------------------------------------------- trace evaluation
def load_fkcc_source(hh_path: str, fn: str = "fkcc"):
    src = open(hh_path).read().split("\n")
    start = None
    for i, line in enumerate(src):
        if f"inline static bool {fn}(" in line:
            start = i
            break
    assert start is not None
    # body runs until the matching "return true;" of fkcc (first one after start at function depth)
    body = []
    for line in src[start:]:
        body.append(line)
        if line.strip() == "return true;":
            break
    return body


def split_trace(body):
    """-> (fk statements as python source, collision-structure lines)"""
    fk_lines, rest = [], []
    in_fk = False
    done_fk = False
    stmt = ""
    for line in body:
        s = line.strip()
        if not done_fk:
            if re.match(r"^(v|y)\[\d+\]\s*=", s) or stmt:
                in_fk = True
                stmt += " " + s
                if s.endswith(";"):
                    fk_lines.append(stmt.strip().rstrip(";"))
                    stmt = ""
                continue
            if in_fk and s.startswith("//") and "environment vs. robot" in s:
                done_fk = True
        if done_fk:
            rest.append(s)
    return fk_lines, rest


def eval_trace(fk_lines, q: np.ndarray, n_y: int, n_v: int):
    """q: [M, dim] float64 -> y: [n_y, M]"""
    m = q.shape[0]
    x = [q[:, j].astype(np.float64) for j in range(q.shape[1])]
    v = [np.zeros(m) for _ in range(n_v)]
    y = [np.zeros(m) for _ in range(n_y)]
    env = {"x": x, "v": v, "y": y, "sin": np.sin, "cos": np.cos}
    for st in fk_lines:
        lhs, rhs = st.split("=", 1)
        val = eval(rhs, {"__builtins__": {}}, env)  # noqa: S307 - arithmetic on generated data only
        if np.isscalar(val):
            val = np.full(m, float(val))
        exec(f"{lhs.strip()} = _val", {"__builtins__": {}}, dict(env, _val=val))  # noqa: S102
    return np.stack(y)


# ----------------------------------------------------------------------------- structure parsing
def parse_structure(rest):
    """-> env_blocks [(link, coarse_y, [fine_y...])], self_blocks [(linkA, linkB, (ca, cb), [(fa, fb)...])]"""
    text = "\n".join(rest)
    env_part, self_part = text.split("robot self-collisions", 1) if "robot self-collisions" in text else (text, "")
    env_blocks = []
    # each env block: "// <link>" comment then calls
    cur = None
    for line in env_part.split("\n"):
        mm = re.match(r"^//\s*([A-Za-z0-9_]+)\s*$", line)
        if mm:
            cur = [mm.group(1), None, []]
            env_blocks.append(cur)
            continue
        mm = re.search(r"sphere_environment_in_collision\(environment,\s*y\[(\d+)\]", line)
        if mm and cur is not None:
            idx = int(mm.group(1))
            if cur[1] is None:
                cur[1] = idx
            else:
                cur[2].append(idx)
    env_blocks = [tuple(b) for b in env_blocks if b[1] is not None]

    self_blocks = []
    cur = None
    joined = re.sub(r"\(\s*\n\s*", "(", self_part)
    for line in joined.split("\n"):
        mm = re.match(r"^//\s*([A-Za-z0-9_]+) vs\. ([A-Za-z0-9_]+)\s*$", line.strip())
        if mm:
            cur = [mm.group(1), mm.group(2), None, []]
            self_blocks.append(cur)
            continue
        mm = re.search(r"sphere_sphere_self_collision<[^>]*>\(\s*y\[(\d+)\],[^)]*?y\[\d+\],\s*y\[\d+\],\s*y\[\d+\],\s*y\[(\d+)\]", line)
        if mm and cur is not None:
            pair = (int(mm.group(1)), int(mm.group(2)))
            if cur[2] is None:
                cur[2] = pair
            else:
                cur[3].append(pair)
    return env_blocks, [tuple(b) for b in self_blocks if b[2] is not None]


# ----------------------------------------------------------------------------- URDF
def rpy_matrix(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def parse_urdf(path):
    root = ET.parse(path).getroot()
    links = {}
    link_order = []
    for ln in root.findall("link"):
        spheres = []
        for col in ln.findall("collision"):
            geom = col.find("geometry")
            sp = geom.find("sphere") if geom is not None else None
            if sp is None:
                continue
            org = col.find("origin")
            xyz = [float(v) for v in (org.get("xyz", "0 0 0") if org is not None else "0 0 0").split()]
            spheres.append((xyz, float(sp.get("radius"))))
        links[ln.get("name")] = spheres
        link_order.append(ln.get("name"))
    joints = []
    for j in root.findall("joint"):
        org = j.find("origin")
        xyz = [float(v) for v in (org.get("xyz", "0 0 0") if org is not None else "0 0 0").split()]
        rpy = [float(v) for v in (org.get("rpy", "0 0 0") if org is not None else "0 0 0").split()]
        ax = j.find("axis")
        axis = [float(v) for v in ax.get("xyz").split()] if ax is not None else [0, 0, 1]
        joints.append(dict(name=j.get("name"), type=j.get("type"), parent=j.find("parent").get("link"),
                           child=j.find("child").get("link"), xyz=xyz, rpy=rpy, axis=axis))
    return links, link_order, joints


def axis_rotation(axis, ang):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(ang) * k + (1 - math.cos(ang)) * (k @ k)


def urdf_fk(joints, actuated, q, link_names):
    """world transform (R, t) of every link for one configuration q (float64)."""
    children = {}
    for j in joints:
        children.setdefault(j["parent"], []).append(j)
    all_children = {j["child"] for j in joints}
    roots = [l for l in link_names if l not in all_children]
    tf = {r: (np.eye(3), np.zeros(3)) for r in roots}
    stack = list(roots)
    while stack:
        p = stack.pop()
        rp, tp = tf[p]
        for j in children.get(p, []):
            ro = rpy_matrix(*j["rpy"])
            r = rp @ ro
            t = tp + rp @ np.asarray(j["xyz"])
            if j["name"] in actuated:
                val = q[actuated.index(j["name"])]
                if j["type"] in ("revolute", "continuous"):
                    r = r @ axis_rotation(j["axis"], val)
                elif j["type"] == "prismatic":
                    t = t + r @ (np.asarray(j["axis"], dtype=np.float64) * val)
            tf[j["child"]] = (r, t)
            stack.append(j["child"])
    return tf


# ----------------------------------------------------------------------------- main
def extract(robot: str):
    hh = f"{REF}/external/vamp/src/impl/vamp/robots/{robot}.hh"
    urdf = f"{REF}/external/vamp/resources/{robot}/{robot}_spherized.urdf"
    src = open(hh).read()

    def const(name, cast=float):
        return cast(re.search(name + r"\s*=\s*([-0-9.e]+)", src).group(1))

    dim = const("dimension", int)
    n_fine = const("n_spheres", int)
    resolution = const("resolution", int)
    jn_block = re.search(r"joint_names = \{(.*?)\};", src, re.S).group(1)
    actuated = re.findall(r'"([^"]+)"', jn_block)
    assert len(actuated) == dim

    def arr(name):
        blk = re.search(name + r"\{(.*?)\};", src, re.S).group(1)
        return [float(v) for v in re.findall(r"[-0-9.e]+", blk)]

    s_m, s_a = arr("s_m"), arr("s_a")

    body = load_fkcc_source(hh)
    n_v = int(re.search(r"FloatVector<rake, 1>, (\d+)> v;", "\n".join(body)).group(1))
    n_y = int(re.search(r"FloatVector<rake, 1>, (\d+)> y;", "\n".join(body)).group(1))
    fk_lines, rest = split_trace(body)
    env_blocks, self_blocks = parse_structure(rest)

    links, link_order, joints = parse_urdf(urdf)

    rng = np.random.default_rng(20260101)
    lo = np.float32(s_a).astype(np.float64)
    hi = (np.float32(s_a) + np.float32(s_m)).astype(np.float64)
    qs = rng.uniform(lo, hi, size=(64, dim))
    y = eval_trace(fk_lines, qs, n_y, n_v)  # [n_y, 64]
    tfs = [urdf_fk(joints, actuated, q, link_order) for q in qs]

    # ---- match each trace fine sphere (index k -> y[4k..4k+3]) to a URDF (link, sphere)
    urdf_spheres = [(ln, i, np.asarray(c), r) for ln in link_order for i, (c, r) in enumerate(links[ln])]
    world = np.zeros((len(urdf_spheres), 64, 3))
    for u, (ln, i, c, r) in enumerate(urdf_spheres):
        for m in range(64):
            rr, tt = tfs[m][ln]
            world[u, m] = rr @ c + tt
    fine = []
    used = set()
    worst = 0.0
    for k in range(n_fine):
        p = y[4 * k:4 * k + 3].T  # [64,3]
        rad = y[4 * k + 3, 0]
        err = np.abs(world - p[None]).max(axis=(1, 2))
        for u in np.argsort(err):
            if u not in used and abs(urdf_spheres[u][3] - rad) < 1e-6:
                break
        assert err[u] < 2e-6, (k, err[u])
        worst = max(worst, err[u])
        used.add(u)
        fine.append(dict(link=urdf_spheres[u][0], local=urdf_spheres[u][2].tolist(), r=float(rad)))
    unused = [urdf_spheres[u][:2] for u in range(len(urdf_spheres)) if u not in used]

    # ---- moving frames: one per actuated joint (+ frame 0 = world for links fixed to the base)
    parent_joint = {j["child"]: j for j in joints}

    def frame_of(link):
        """(frame_id, R_off, t_off): link pose = frame pose * offset (constant)"""
        r_off, t_off = np.eye(3), np.zeros(3)
        cur = link
        while cur in parent_joint:
            j = parent_joint[cur]
            if j["name"] in actuated:
                return actuated.index(j["name"]) + 1, r_off, t_off
            ro = rpy_matrix(*j["rpy"])
            t_off = ro @ t_off + np.asarray(j["xyz"])
            r_off = ro @ r_off
            cur = j["parent"]
        return 0, r_off, t_off

    frames = [dict(id=0, parent=-1, joint=None)]
    for idx, jn in enumerate(actuated):
        j = next(jj for jj in joints if jj["name"] == jn)
        pf, r_off, t_off = frame_of(j["parent"])
        ro = rpy_matrix(*j["rpy"])
        r_c = r_off @ ro
        t_c = t_off + r_off @ np.asarray(j["xyz"])
        frames.append(dict(id=idx + 1, parent=pf, joint=jn, type=j["type"], axis=j["axis"], q_index=idx,
                           R=r_c.tolist(), t=t_c.tolist()))

    # ---- links as the reference groups them (env blocks), coarse spheres solved in frame coordinates
    link_frame = {}
    model_links = []
    fine_by_y = {4 * k: k for k in range(n_fine)}
    for (lname, cy, fys) in env_blocks:
        fid, r_off, t_off = frame_of(lname)
        link_frame[lname] = fid
        # frame world pose per sample
        locs = []
        for m in range(64):
            rr, tt = tfs[m][lname]
            # frame pose = link pose * offset^-1
            rf = rr @ r_off.T
            tf_ = tt - rf @ t_off
            pw = y[cy:cy + 3, m]
            locs.append(rf.T @ (pw - tf_))
        locs = np.asarray(locs)
        assert locs.std(axis=0).max() < 1e-7, (lname, locs.std(axis=0))
        fine_ids = [fine_by_y[f] for f in fys]
        fl = []
        for k in fine_ids:
            assert fine[k]["link"] == lname or frame_of(fine[k]["link"])[0] == fid
            _, r2, t2 = frame_of(fine[k]["link"])
            fl.append(dict(id=k, c=(r2 @ np.asarray(fine[k]["local"]) + t2).tolist(), r=fine[k]["r"]))
        model_links.append(dict(name=lname, frame=fid, coarse=dict(c=locs.mean(axis=0).tolist(), r=float(y[cy + 3, 0])),
                                fine=fl))

    name_to_link = {l["name"]: i for i, l in enumerate(model_links)}
    coarse_y = {cy: lname for (lname, cy, _) in env_blocks}
    pairs = []
    for (a, b, (ca, cb), fps) in self_blocks:
        assert coarse_y[ca] == a and coarse_y[cb] == b
        fa = [f["id"] for f in model_links[name_to_link[a]]["fine"]]
        fb = [f["id"] for f in model_links[name_to_link[b]]["fine"]]
        want = {(4 * i, 4 * j) for i in fa for j in fb}
        assert set(fps) == want, (a, b, len(fps), len(want))
        pairs.append([name_to_link[a], name_to_link[b]])

    # ---- bounding property of the hierarchy: slack = r_coarse - (|c_f - c_c| + r_f)
    slack = min(l["coarse"]["r"] - (np.linalg.norm(np.asarray(f["c"]) - np.asarray(l["coarse"]["c"])) + f["r"])
                for l in model_links for f in l["fine"])

    # ---- attachments (Robot::fkcc_attach): the end-effector frame the attachment is posed in and the links tested against it.
    # The trace of fkcc_attach is evaluated like fkcc's; `set_attachment_pose(environment, to_isometry(&y[N]))` names the 12
    # outputs holding the pose (translation, then the rotation column by column: vector/math.hh:39-51); it must be a constant
    # transform away from one of the moving frames. "// Attachment vs. <link>" blocks list bounding sphere, then fine spheres.
    abody = load_fkcc_source(hh, "fkcc_attach")
    a_nv = int(re.search(r"FloatVector<rake, 1>, (\d+)> v;", "\n".join(abody)).group(1))
    a_ny = int(re.search(r"FloatVector<rake, 1>, (\d+)> y;", "\n".join(abody)).group(1))
    a_fk, a_rest = split_trace(abody)
    ya = eval_trace(a_fk, qs, a_ny, a_nv)
    assert np.abs(ya[:n_y] - y).max() < 1e-12, "fkcc_attach must lay out its spheres like fkcc"
    ee_off = int(re.search(r"set_attachment_pose\(environment, to_isometry\(&y\[(\d+)\]\)\)", "\n".join(a_rest)).group(1))
    ee_t = ya[ee_off:ee_off + 3].T                                        # [64, 3]
    ee_r = ya[ee_off + 3:ee_off + 12].T.reshape(64, 3, 3).transpose(0, 2, 1)  # column-major -> [64][row][col]
    ee = None
    for idx in reversed(range(len(actuated))):
        j = next(jj for jj in joints if jj["name"] == actuated[idx])
        loc_r, loc_t = [], []
        for m in range(64):
            rf, tf_ = tfs[m][j["child"]]
            loc_r.append(rf.T @ ee_r[m])
            loc_t.append(rf.T @ (ee_t[m] - tf_))
        loc_r, loc_t = np.asarray(loc_r), np.asarray(loc_t)
        if loc_r.std(axis=0).max() < 1e-7 and loc_t.std(axis=0).max() < 1e-7:
            ee = dict(frame=idx + 1, R=loc_r.mean(axis=0).tolist(), t=loc_t.mean(axis=0).tolist())
            break
    assert ee is not None, "end-effector pose is not rigidly attached to any moving frame"
    att_links = []
    cur = None
    for line in a_rest:
        mm = re.match(r"^//\s*Attachment vs\. ([A-Za-z0-9_]+)\s*$", line)
        if mm:
            cur = [mm.group(1), None, []]
            att_links.append(cur)
            continue
        mm = re.search(r"attachment_sphere_collision<[^>]*>\(\s*environment,\s*y\[(\d+)\]", line)
        if mm and cur is not None:
            if cur[1] is None:
                cur[1] = int(mm.group(1))
            else:
                cur[2].append(int(mm.group(1)))
    att = []
    for (lname, cy, fys) in att_links:
        li = name_to_link[lname]
        assert coarse_y[cy] == lname, (lname, cy)
        assert sorted(fys) == sorted(4 * f["id"] for f in model_links[li]["fine"]), lname
        att.append(li)
    attachment = dict(ee_frame=ee["frame"], R=ee["R"], t=ee["t"], links=att, link_names=[a[0] for a in att_links])

    model = dict(        name=robot, dimension=dim, n_fine=n_fine, resolution=resolution,
        s_a=s_a, s_m=s_m, joint_names=actuated,
        min_radius=const("min_radius"), max_radius=const("max_radius"),
        frames=frames, links=model_links, self_pairs=pairs, attachment=attachment,
        provenance=dict(urdf=os.path.relpath(urdf, REF), trace=os.path.relpath(hh, REF),
                        fine_match_max_err_m=float(worst), urdf_spheres_not_in_reference=unused,
                        min_bounding_slack_m=float(slack),
                        n_self_fine_pairs=sum(len(b[3]) for b in self_blocks)),
    )
    return model, (fk_lines, n_y, n_v)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("robots", nargs="*", default=["panda", "fetch"])
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "ompl_b200", "robots"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for r in a.robots:
        model, _ = extract(r)
        p = os.path.join(a.out, f"{r}.json")
        with open(p, "w") as f:
            json.dump(model, f, indent=1)
        pv = model["provenance"]
        print(f"{r}: dim={model['dimension']} fine={model['n_fine']} links={len(model['links'])} "
              f"pairs={len(model['self_pairs'])} fine_pairs={pv['n_self_fine_pairs']} "
              f"match_err={pv['fine_match_max_err_m']:.2e} slack={pv['min_bounding_slack_m']:.3e} "
              f"unused={pv['urdf_spheres_not_in_reference']}", file=sys.stderr)
