"""oracle/map_flatten.py -- TEST INFRASTRUCTURE ONLY.

Object-graph restatement of the graph construction of Optimizer::LocalGPBA (src/Optimizer.cc:718-1211), its tail
(:1349-1430) and Optimizer::BundleAdjustment (:85-315): MultiKeyFrame / MapPoint objects with mPrevKF / mNextKF pointers,
per-point observation maps, mnBALocalForKF / mnBAFixedForKF stamps -- written as the reference writes it (pointer chasing,
one loop per reference loop), with one stated deviation: a point's observations are visited in ascending keyframe id
(the reference's std::map is keyed by keyframe pointer, i.e. allocation order).  The product (csrc/gpba_map.cc) keeps a
slot-addressed SoA mirror instead; tests/test_map_mirror.py checks that both produce identical gpba_problem arrays.
PARITY UNPINNED against the reference binary (it cannot be built here).
"""
import math

import numpy as np


class KF:
    def __init__(self, id, prev, pose, vel, time, cam_time):
        self.mnId, self.mPrevKF, self.mNextKF = id, prev, None
        self.pose, self.vel, self.time, self.cam_time = np.array(pose, float), np.array(vel, float), float(time), np.array(cam_time, float)
        self.bad = False
        self.matches = {}            # (cam, point id) -> MapPoint, insertion order = GetMapPointMatches order
        self.mnBALocalForKF = self.mnBAFixedForKF = -1
        self.mConnectedKeyFrameWeights = {}      # KF -> weight
        self.mvpOrderedConnectedKeyFrames, self.mvOrderedWeights = [], []

    # src/KeyFrame.cc:250-287
    def AddConnection(self, pKF, weight):
        if pKF not in self.mConnectedKeyFrameWeights or self.mConnectedKeyFrameWeights[pKF] != weight:
            self.mConnectedKeyFrameWeights[pKF] = weight
        else:
            return
        self.UpdateBestCovisibles()

    def UpdateBestCovisibles(self):
        vPairs = sorted(((w, k.mnId, k) for k, w in self.mConnectedKeyFrameWeights.items()), key=lambda t: (t[0], t[1]))
        lKFs, lWs = [], []
        for w, _, k in vPairs:
            if not k.bad:
                lKFs.insert(0, k); lWs.insert(0, w)
        self.mvpOrderedConnectedKeyFrames, self.mvOrderedWeights = lKFs, lWs

    def EraseConnection(self, pKF):
        if pKF in self.mConnectedKeyFrameWeights:
            del self.mConnectedKeyFrameWeights[pKF]
            self.UpdateBestCovisibles()


class MP:
    def __init__(self, id, xyz):
        self.mnId, self.xyz, self.bad = id, np.array(xyz, float), False
        self.obs = {}                # KF id -> {cam: (u, v, ur, w, close)}
        self.mnBALocalForKF = -1


class RefMap:
    def __init__(self, cam_intr, cam_Tbc, bf, qc):
        self.cam_intr, self.cam_Tbc, self.bf, self.qc = np.array(cam_intr, float).reshape(-1, 4), np.array(cam_Tbc, float).reshape(-1, 7), float(bf), np.array(qc, float)
        self.n_cam = len(self.cam_intr)
        self.kfs, self.pts = {}, {}
        self.stamp = 0

    # ---- mutation
    def add_keyframe(self, id, prev_id, pose, vel, time, cam_time):
        prev = self.kfs[prev_id] if prev_id >= 0 else None
        k = KF(id, prev, pose, vel, time, cam_time)
        if prev is not None:
            prev.mNextKF = k
        self.kfs[id] = k

    def set_keyframe_state(self, id, pose, vel=None):
        self.kfs[id].pose = np.array(pose, float)
        if vel is not None:
            self.kfs[id].vel = np.array(vel, float)

    def set_keyframe_bad(self, id):
        k = self.kfs[id]
        if k.bad:
            return
        if k.mPrevKF is not None and k.mNextKF is not None:
            k.mNextKF.mPrevKF = k.mPrevKF
            k.mPrevKF.mNextKF = k.mNextKF
            k.mNextKF = k.mPrevKF = None
        for other in list(k.mConnectedKeyFrameWeights):       # SetBadFlag (KeyFrame.cc:663-666)
            other.EraseConnection(k)
        k.mConnectedKeyFrameWeights.clear(); k.mvpOrderedConnectedKeyFrames = []; k.mvOrderedWeights = []
        for (cam, pid), p in list(k.matches.items()):
            del p.obs[id][cam]
            if not p.obs[id]:
                del p.obs[id]
        k.matches.clear()
        k.bad = True

    def add_point(self, id, xyz):
        self.pts[id] = MP(id, xyz)

    def set_point(self, id, xyz):
        self.pts[id].xyz = np.array(xyz, float)

    def set_point_bad(self, id):
        p = self.pts[id]
        if p.bad:
            return
        for kid, cams in p.obs.items():
            for cam in cams:
                del self.kfs[kid].matches[(cam, id)]
        p.obs.clear()
        p.bad = True

    def add_observation(self, kf, cam, pt, u, v, ur, w, close):
        k, p = self.kfs[kf], self.pts[pt]
        p.obs.setdefault(kf, {})[cam] = (float(u), float(v), float(ur), float(w), 1 if close else 0)
        if (cam, pt) not in k.matches:
            k.matches[(cam, pt)] = p

    def erase_observation(self, kf, cam, pt):
        p = self.pts[pt]
        del p.obs[kf][cam]
        if not p.obs[kf]:
            del p.obs[kf]
        del self.kfs[kf].matches[(cam, pt)]

    # MultiKeyFrame::UpdateConnections (src/KeyFrame.cc:455-550); the pointer-keyed maps are walked by keyframe id here
    def update_connections(self, kf_id):
        this = self.kfs[kf_id]
        KFcounter = {}
        for p in this.matches.values():                      # vpMP = mvpMapPoints: one entry per keypoint
            if p.bad:
                continue
            for oid in p.obs:                                # GetObservations(): one entry per keyframe
                o = self.kfs[oid]
                if o.mnId == this.mnId or o.bad:
                    continue
                KFcounter[o] = KFcounter.get(o, 0) + 1
        if not KFcounter:
            return
        nmax, pKFmax, th, vPairs = 0, None, 15, []
        for o in sorted(KFcounter, key=lambda k: k.mnId):
            w = KFcounter[o]
            if w > nmax:
                nmax, pKFmax = w, o
            if w >= th:
                vPairs.append((w, o.mnId, o))
                o.AddConnection(this, w)
        if not vPairs:
            vPairs.append((nmax, pKFmax.mnId, pKFmax))
            pKFmax.AddConnection(this, nmax)
        vPairs.sort(key=lambda t: (t[0], t[1]))
        this.mConnectedKeyFrameWeights = dict(KFcounter)
        this.mvpOrderedConnectedKeyFrames = [t[2] for t in reversed(vPairs)]
        this.mvOrderedWeights = [t[0] for t in reversed(vPairs)]

    def covisibles(self, kf_id):
        k = self.kfs[kf_id]
        return [o.mnId for o in k.mvpOrderedConnectedKeyFrames], list(k.mvOrderedWeights)

    def n_alive(self):
        return sum(1 for k in self.kfs.values() if not k.bad)

    # ---- flattening
    def _emit(self, out, p, pt_index, in_graph, index_of):
        n_cam = self.n_cam
        for kid in sorted(p.obs):
            k = self.kfs[kid]
            if not in_graph(k):
                continue
            idxs = p.obs[kid]
            for c in range(n_cam - 1):
                if c not in idxs:
                    continue
                prev = k.mPrevKF
                if prev is None or not in_graph(prev):
                    continue
                out["edges"].append((index_of[prev.mnId], index_of[kid], c, k.cam_time[c], idxs[c], pt_index, kid, p.mnId))
                out["cam_obs"][c] += 1
            c = n_cam - 1
            if c in idxs:
                out["edges"].append((-1, index_of[kid], c, k.time, idxs[c], pt_index, kid, p.mnId))

    def _pack(self, members, roles, pts, out, velp, priors, huber_prior, lambda_init):
        # records numbered by (keyframe index, camera) among those in use (the product numbers them the same way; the
        # reference has no records, only edges)
        keys = sorted({(k2, c) for (_, k2, c, _, _, _, _, _) in out["edges"]})
        rec_of = {key: i for i, key in enumerate(keys)}
        rec = [None] * len(keys)
        obs = dict(u=[], v=[], ur=[], w=[], rec=[], pt=[], fl=[], kf=[], cam=[], pid=[])
        for (k1, k2, c, t, o, pi, kid, pid) in out["edges"]:
            key = (k2, c)
            rec[rec_of[key]] = (k1, k2, c, t)
            u, v, ur, w, close = o
            stereo = c == self.n_cam - 1 and ur >= 0
            obs["u"].append(u); obs["v"].append(v); obs["ur"].append(ur if stereo else -1.0); obs["w"].append(w)
            obs["rec"].append(rec_of[key]); obs["pt"].append(pi); obs["fl"].append(1 if close else 0)
            obs["kf"].append(kid); obs["cam"].append(c); obs["pid"].append(pid)
        any_stereo = any(x >= 0 for x in obs["ur"])
        return dict(
            cam_intr=self.cam_intr, cam_Tbc=self.cam_Tbc, bf=self.bf, qc=self.qc,
            kf_pose=np.array([k.pose for k in members]).reshape(-1, 7), kf_vel=np.array([k.vel for k in members]).reshape(-1, 6),
            kf_time=np.array([k.time for k in members]), kf_fixed=np.array([1 if r == 2 else 0 for r in roles], np.uint8),
            kf_id=np.array([k.mnId for k in members], np.int64), kf_role=np.array(roles, np.int32),
            pt_xyz=np.array([p.xyz for p in pts]).reshape(-1, 3), pt_id=np.array([p.mnId for p in pts], np.int64),
            rec_kf1=np.array([r[0] for r in rec], np.int32), rec_kf2=np.array([r[1] for r in rec], np.int32),
            rec_cam=np.array([r[2] for r in rec], np.int32), rec_t=np.array([r[3] for r in rec], float),
            obs_u=np.array(obs["u"], float), obs_v=np.array(obs["v"], float), obs_ur=np.array(obs["ur"], float) if any_stereo else None,
            obs_inv_sigma2=np.array(obs["w"], float), obs_rec=np.array(obs["rec"], np.int32), obs_pt=np.array(obs["pt"], np.int32),
            obs_flags=np.array(obs["fl"], np.uint8), obs_kf=np.array(obs["kf"], np.int64), obs_cam=np.array(obs["cam"], np.int32),
            obs_pt_id=np.array(obs["pid"], np.int64),
            prior_kf1=np.array([a for a, _ in priors], np.int32), prior_kf2=np.array([b for _, b in priors], np.int32),
            velp_kf=np.array(velp, np.int32), cam_obs=np.array(out["cam_obs"], np.int32),
            huber_mono=float(np.float32(math.sqrt(5.991))), huber_stereo=float(np.float32(math.sqrt(7.815))),
            huber_prior=huber_prior, lambda_init=lambda_init)

    def local_window(self, kf_id, large=False, covisible=()):
        if covisible is None:
            covisible = [o.mnId for o in self.kfs[kf_id].mvpOrderedConnectedKeyFrames]
        self.stamp += 1
        S = self.stamp
        pKF = self.kfs[kf_id]
        maxOpt = 25 if large else 10
        Nd = min(self.n_alive() - 2, maxOpt)
        opt = [pKF]
        pKF.mnBALocalForKF = S
        for _ in range(1, Nd):
            if opt[-1].mPrevKF is not None:
                opt.append(opt[-1].mPrevKF)
                opt[-1].mnBALocalForKF = S
            else:
                break
        lpts = []

        def add_points(k):
            for p in k.matches.values():
                if not p.bad and p.mnBALocalForKF != S:
                    lpts.append(p)
                    p.mnBALocalForKF = S
        for k in opt:
            add_points(k)
        fixed = []
        if opt[-1].mPrevKF is not None:
            fixed.append(opt[-1].mPrevKF)
            opt[-1].mPrevKF.mnBAFixedForKF = S
        else:
            opt[-1].mnBALocalForKF = -1
            opt[-1].mnBAFixedForKF = S
            fixed.append(opt[-1])
            opt.pop()
        vis = []
        for cid in covisible:
            if len(vis) > 0:
                break
            k = self.kfs.get(cid)
            if k is None:
                continue
            if k.mnBALocalForKF == S or k.mnBAFixedForKF == S:
                continue
            k.mnBALocalForKF = S
            if not k.bad:
                vis.append(k)
                add_points(k)
        for p in lpts:
            for kid in sorted(p.obs):
                k = self.kfs[kid]
                if k.mnBALocalForKF != S and k.mnBAFixedForKF != S:
                    k.mnBAFixedForKF = S
                    if not k.bad:
                        fixed.append(k)
                        break
            if len(fixed) >= 50:
                break
        members = sorted([(k, 0) for k in opt] + [(k, 1) for k in vis] + [(k, 2) for k in fixed], key=lambda t: t[0].mnId)
        index_of = {k.mnId: i for i, (k, _) in enumerate(members)}
        velp = [index_of[k.mnId] for k in opt]
        priors = [(index_of[opt[i].mnId], index_of[opt[i - 1].mnId]) for i in range(len(opt) - 1, 0, -1)]
        out = dict(edges=[], cam_obs=[0] * self.n_cam)
        in_graph = lambda k: (not k.bad) and (k.mnBALocalForKF == S or k.mnBAFixedForKF == S)
        for i, p in enumerate(lpts):
            self._emit(out, p, i, in_graph, index_of)
        return self._pack([k for k, _ in members], [r for _, r in members], lpts, out, velp, priors, 0.0, 1e-2 if large else 1.0)

    def global_window(self, init_kf_id):
        members = sorted((k for k in self.kfs.values() if not k.bad), key=lambda k: k.mnId)
        index_of = {k.mnId: i for i, k in enumerate(members)}
        roles = [2 if k.mnId == init_kf_id else 0 for k in members]
        velp, priors = [], []
        for k in members:
            velp.append(index_of[k.mnId])
            if k.mPrevKF is not None and not k.mPrevKF.bad:
                priors.append((index_of[k.mPrevKF.mnId], index_of[k.mnId]))
        out = dict(edges=[], cam_obs=[0] * self.n_cam)
        pts = []
        in_graph = lambda k: not k.bad
        for p in sorted((p for p in self.pts.values() if not p.bad), key=lambda p: p.mnId):
            if not any(not self.kfs[kid].bad for kid in p.obs):
                continue
            self._emit(out, p, len(pts), in_graph, index_of)
            pts.append(p)
        return self._pack(members, roles, pts, out, velp, priors, 21.026, 1e-5)
