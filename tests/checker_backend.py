"""TEST INFRASTRUCTURE: a stand-in for ompl_b200.roadmap.EngineBackend whose four phases are computed by the CPU
checker (oracle/port.py) on CPU tensors. It lets the world_size-2 gloo tests exercise the sharding, collectives and
assembly logic of RoadmapBuilder without a GPU, and it is the ground truth the GPU roadmap is compared with."""
import numpy as np
import torch

from oracle import port


class CheckerBackend:
    def __init__(self, port_env, robot: str, nthreads: int = 4):
        self.env, self.robot, self.nthreads = port_env, robot, nthreads
        self.dim = port.robot_info(robot)["dim"]
        self.device = torch.device("cpu")

    def sample(self, seed, first, count, out):
        cand = port.prm_candidates(self.robot, seed, first, count)
        ok = self.env.validate_states(self.robot, cand, nthreads=self.nthreads) == 1
        rows = torch.from_numpy(cand[ok])
        m = min(rows.shape[0], out.shape[0])
        out[:m] = rows[:m]
        return int(rows.shape[0])

    def neighbors_part(self, states, k, part, parts, idx, dst):
        n = states.shape[0]
        per = -(-n // parts)
        b, e = min(per * part, n), min(per * (part + 1), n)
        if e > b:
            i, d = self.neighbors(states, b, e - b, k)
            idx[b:e] = i
            dst[b:e] = d

    def neighbors(self, states, q_first, q_count, k):
        db = states.numpy().astype(np.float64)
        idx, dst = port.knn(db, db[q_first:q_first + q_count], k + 1, nthreads=self.nthreads)
        oi = np.full((q_count, k), -1, np.int64)
        od = np.full((q_count, k), np.inf)
        for r in range(q_count):
            keep = idx[r] != (q_first + r)
            if keep.all():
                keep[-1] = False
            oi[r] = idx[r][keep][:k]
            od[r] = dst[r][keep][:k]
        return torch.from_numpy(oi), torch.from_numpy(od)

    def edges(self, nbr, dist_local, q_first, q_count):
        e, w = port.prm_edge_list(nbr.numpy(), dist_local.numpy(), q_first, q_count)
        return torch.from_numpy(e), torch.from_numpy(w)

    def validate_edges(self, states, edges, resolution):
        s = states.numpy()
        e = edges.numpy()
        if len(e) == 0:
            return torch.zeros(0, dtype=torch.uint8)
        return torch.from_numpy(self.env.validate_motions(self.robot, s[e[:, 0]], s[e[:, 1]], resolution=resolution, nthreads=self.nthreads))
