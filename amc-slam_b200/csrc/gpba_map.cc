// gpba_map.cc -- host-side SoA mirror of the AMC-SLAM map and the flattening of LocalGPBA / BundleAdjustment graphs
// into gpba_problem arrays (include/gpba_map.h).  No device work in this file.
//
// Follows the graph construction of Optimizer::LocalGPBA (src/Optimizer.cc:718-1211), its tail (:1349-1430) and
// Optimizer::BundleAdjustment (:85-315).  Storage: keyframes, points and observations live in flat vectors addressed by
// slot; a point keeps its observation slots sorted by (keyframe id, camera) -- the iteration order of the reference's
// per-point std::map + vector<int> with ids in place of pointers -- and a keyframe keeps its slots in insertion order
// (GetMapPointMatches order).  Window membership uses stamps (mnBALocalForKF / mnBAFixedForKF), so selecting a window
// touches only the keyframes and points that end up in it.
#include "../../include/gpba_map.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <thread>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace {

struct Obs {                // pool entry: identity + liveness (what the keyframe lists and windows refer to)
  int kf = -1, pt = -1, cam = 0;
  bool alive = false;
};
struct PObs {               // payload, stored contiguously in the point's own list: flattening a point reads one array
  int slot, kf, cam;
  uint8_t close_flag;
  double u, v, ur, w;
};
struct Kf {
  int64_t id = 0;
  int prev = -1, next = -1;
  double pose[7], vel[6], time = 0;
  std::vector<double> cam_time;
  bool bad = false;
  std::vector<int> obs;   // observation slots in insertion order; dead slots are skipped and compacted lazily
  int n_dead = 0;
  int64_t local_stamp = -1, fixed_stamp = -1;
  int widx = -1;          // index in the window being built
  std::unordered_map<int, int> conn;            // mConnectedKeyFrameWeights: keyframe slot -> shared map points
  std::vector<std::pair<int, int> > ordered;    // mvpOrderedConnectedKeyFrames / mvOrderedWeights: (slot, weight)
};
struct Pt {
  int64_t id = 0;
  double xyz[3];
  bool bad = false;
  std::vector<PObs> obs;  // observations sorted by (keyframe id, camera)
  int64_t stamp = -1;
};

}  // namespace

struct gpba_map {
  int n_cam = 0;
  std::vector<double> cam_intr, cam_Tbc;
  double bf = 0, qc[6];
  std::vector<Kf> kfs;
  std::vector<Pt> pts;
  std::vector<Obs> obs;
  std::vector<int> free_obs;
  std::unordered_map<int64_t, int> kf_of, pt_of;
  int64_t stamp = 0, n_kf_alive = 0, n_pt_alive = 0, n_obs_alive = 0;

  int find_obs(const Pt& p, int64_t kf_id, int cam, size_t* pos) const {
    auto it = std::lower_bound(p.obs.begin(), p.obs.end(), std::make_pair(kf_id, cam), [&](const PObs& o, const std::pair<int64_t, int>& key) {
      const int64_t id = kfs[o.kf].id;
      return id < key.first || (id == key.first && o.cam < key.second);
    });
    if (pos) *pos = (size_t)(it - p.obs.begin());
    if (it != p.obs.end() && kfs[it->kf].id == kf_id && it->cam == cam) return it->slot;
    return -1;
  }
  void kill_obs(int s, bool from_point, bool may_compact = true) {
    Obs& o = obs[s];
    if (!o.alive) return;
    o.alive = false;
    --n_obs_alive;
    if (from_point) {
      Pt& p = pts[o.pt];
      size_t pos;
      if (find_obs(p, kfs[o.kf].id, o.cam, &pos) == s) p.obs.erase(p.obs.begin() + pos);
    }
    Kf& k = kfs[o.kf];
    if (++k.n_dead > 64 && (size_t)k.n_dead * 2 > k.obs.size() && may_compact) compact(k);
  }
  // MultiKeyFrame::UpdateBestCovisibles (KeyFrame.cc:265-287): all connected, not bad, by (weight, id) descending
  void update_best_covisibles(Kf& k) {
    std::vector<std::pair<std::pair<int, int64_t>, int> > v;
    v.reserve(k.conn.size());
    for (const auto& c : k.conn) if (!kfs[c.first].bad) v.push_back({{c.second, kfs[c.first].id}, c.first});
    std::sort(v.begin(), v.end());
    k.ordered.clear();
    for (size_t i = v.size(); i-- > 0;) k.ordered.push_back({v[i].second, v[i].first.first});
  }
  void add_connection(int to, int from, int weight) {   // kfs[to].AddConnection(kfs[from], weight) (:250-263)
    Kf& k = kfs[to];
    auto it = k.conn.find(from);
    if (it != k.conn.end() && it->second == weight) return;
    k.conn[from] = weight;
    update_best_covisibles(k);
  }
  void compact(Kf& k) {
    size_t w = 0;
    for (size_t i = 0; i < k.obs.size(); ++i) {
      if (obs[k.obs[i]].alive) k.obs[w++] = k.obs[i];
      else free_obs.push_back(k.obs[i]);
    }
    k.obs.resize(w);
    k.n_dead = 0;
  }
};

struct gpba_window {
  bool local = true, large = false;
  int iterations = 10;
  gpba_problem P;
  std::vector<double> cam_intr, cam_Tbc, kf_pose, kf_vel, kf_time, pt_xyz, rec_t, obs_u, obs_v, obs_ur, obs_w;
  std::vector<uint8_t> kf_fixed, obs_flags;
  std::vector<int32_t> rec_kf1, rec_kf2, rec_cam, obs_rec, obs_pt, prior_kf1, prior_kf2, velp_kf, kf_role, cam_obs;
  std::vector<int> kf_slot, pt_slot, obs_slot;
  std::vector<int64_t> kf_id, pt_id;
  bool any_stereo = false;
};

namespace {

// Windows are recycled through a small process-wide free list: a local-BA window is ~10 MB of arrays, and handing fresh
// pages to every call costs more (first-touch page faults) than filling them.
std::mutex g_spare_mutex;
std::vector<gpba_window*> g_spare;
gpba_window* new_window() {
  gpba_window* w = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_spare_mutex);
    if (!g_spare.empty()) { w = g_spare.back(); g_spare.pop_back(); }
  }
  if (!w) return new gpba_window();
  w->local = true; w->large = false; w->iterations = 10; w->any_stereo = false;
  // the per-observation arrays keep their old length: the emitter resizes them to the new count and overwrites every element,
  // and a clear() here would make that resize zero-fill ~50 B per observation on one thread first (1 ms at C2)
  for (auto* v : {&w->cam_intr, &w->cam_Tbc, &w->kf_pose, &w->kf_vel, &w->kf_time, &w->pt_xyz, &w->rec_t}) v->clear();
  for (auto* v : {&w->kf_fixed}) v->clear();
  for (auto* v : {&w->rec_kf1, &w->rec_kf2, &w->rec_cam, &w->prior_kf1, &w->prior_kf2, &w->velp_kf, &w->kf_role, &w->cam_obs}) v->clear();
  for (auto* v : {&w->kf_slot, &w->pt_slot}) v->clear();
  for (auto* v : {&w->kf_id, &w->pt_id}) v->clear();
  return w;
}

thread_local std::string g_map_err;
int fail(const char* msg) { g_map_err = msg; return GPBA_ERR_INVALID; }

inline double f32(double x) { return (double)(float)x; }

// Emits the reprojection edges of one point (Optimizer.cc:1037-1210 local, :151-240 global).
// Threads for the per-point passes (GPBA_MAP_THREADS overrides; 1 below 32k observations).
int map_threads(size_t work) {
  static const int env = [] { const char* e = std::getenv("GPBA_MAP_THREADS"); return e ? std::atoi(e) : 0; }();
  if (work < 32768) return 1;
  int t = env > 0 ? env : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  return (int)std::min<size_t>((size_t)t, work / 16384 + 1);
}
template <class F>
void parallel_chunks(size_t n, int threads, F f) {   // f(begin, end, thread)
  if (threads <= 1) { f((size_t)0, n, 0); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back([=] { f(n * t / threads, n * (t + 1) / threads, t); });
  for (auto& th : pool) th.join();
}

// Emits the reprojection edges of the window's points (Optimizer.cc:1037-1210 local, :151-240 global) in two passes over
// the per-point observation lists: count (+ mark the (keyframe, camera) records in use), then fill at the prefix offsets.
// Both passes are independent per point and run on a few host threads; records are numbered by (keyframe, camera).
struct Emitter {
  gpba_map& m;
  gpba_window& w;
  Emitter(gpba_map& m_, gpba_window& w_) : m(m_), w(w_) {}

  template <class InGraph>
  void run(const std::vector<int>& pslots, size_t n_kf_w, InGraph in_graph) {
    const size_t np = pslots.size();
    const int n_cam = m.n_cam;
    size_t work = 0;
    for (int ps : pslots) work += m.pts[ps].obs.size();
    const int T = map_threads(work);
    static const bool timing = std::getenv("GPBA_MAP_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!timing) return;
      auto t1 = std::chrono::steady_clock::now();
      std::fprintf(stderr, "    [emit] %-10s %.3f ms (T=%d)\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count(), T);
      t0 = t1;
    };
    std::vector<int64_t> offset(np + 1, 0);
    std::vector<std::vector<uint8_t>> used(T, std::vector<uint8_t>(n_kf_w * n_cam, 0));
    // walks one point; calls visit(o, kw, pw) for every edge
    auto walk = [&](const Pt& p, auto&& visit) {
      size_t i = 0;
      const size_t no = p.obs.size();
      while (i < no) {
        const int kf = p.obs[i].kf;
        size_t j = i;
        while (j < no && p.obs[j].kf == kf) ++j;   // [i, j): this keyframe's cameras, ascending
        const int kw = in_graph(kf);
        if (kw >= 0) {
          const Kf& K = m.kfs[kf];
          const int pw = K.prev >= 0 ? in_graph(K.prev) : -1;
          for (size_t q = i; q < j; ++q) {
            const PObs& o = p.obs[q];
            if (o.cam < n_cam - 1 && pw < 0) continue;   // no previous keyframe in the graph: no GP edge (:1112-1119, :169)
            visit(o, K, kw, pw);
          }
        }
        i = j;
      }
    };
    parallel_chunks(np, T, [&](size_t b, size_t e, int t) {
      uint8_t* u = used[t].data();
      for (size_t i = b; i < e; ++i) {
        int64_t c = 0;
        walk(m.pts[pslots[i]], [&](const PObs& o, const Kf&, int kw, int) { ++c; u[(size_t)kw * n_cam + o.cam] = 1; });
        offset[i + 1] = c;
      }
    });
    lap("count");
    for (size_t i = 0; i < np; ++i) offset[i + 1] += offset[i];
    const size_t n = (size_t)offset[np];
    std::vector<int> rec_of(n_kf_w * n_cam, -1);
    for (size_t kw = 0; kw < n_kf_w; ++kw)
      for (int c = 0; c < n_cam; ++c) {
        bool any = false;
        for (int t = 0; t < T; ++t) any |= used[t][kw * n_cam + c] != 0;
        if (!any) continue;
        const Kf& K = m.kfs[w.kf_slot[kw]];
        const bool gp = c < n_cam - 1;
        rec_of[kw * n_cam + c] = (int)w.rec_kf2.size();
        w.rec_kf1.push_back(gp ? m.kfs[K.prev].widx : -1); w.rec_kf2.push_back((int)kw); w.rec_cam.push_back(c);
        w.rec_t.push_back(gp ? K.cam_time[c] : K.time);
      }
    w.obs_u.resize(n); w.obs_v.resize(n); w.obs_ur.resize(n); w.obs_w.resize(n);
    w.obs_rec.resize(n); w.obs_pt.resize(n); w.obs_flags.resize(n); w.obs_slot.resize(n);
    lap("resize");
    std::vector<std::vector<int32_t>> cam_obs(T, std::vector<int32_t>(n_cam, 0));
    std::vector<uint8_t> stereo_seen(T, 0);
    parallel_chunks(np, T, [&](size_t b, size_t e, int t) {
      // per-thread tallies live on the thread's stack until the end: the shared arrays sit in one cache line, and a write
      // per observation from every thread made this pass 4x slower than the copies themselves
      std::vector<int32_t> cam_local(n_cam, 0);
      bool stereo_local = false;
      double *ou = w.obs_u.data(), *ov = w.obs_v.data(), *our = w.obs_ur.data(), *ow = w.obs_w.data();
      int32_t *orec = w.obs_rec.data(), *opt = w.obs_pt.data();
      auto* oflags = w.obs_flags.data();
      auto* oslot = w.obs_slot.data();
      const int* rec = rec_of.data();
      for (size_t i = b; i < e; ++i) {
        size_t k = (size_t)offset[i];
        walk(m.pts[pslots[i]], [&](const PObs& o, const Kf&, int kw, int) {
          const bool stereo = o.cam == n_cam - 1 && o.ur >= 0;
          ou[k] = o.u; ov[k] = o.v; our[k] = stereo ? o.ur : -1.0; ow[k] = o.w;
          orec[k] = rec[(size_t)kw * n_cam + o.cam]; opt[k] = (int32_t)i;
          oflags[k] = o.close_flag ? GPBA_OBS_CLOSE : 0; oslot[k] = o.slot;
          if (o.cam < n_cam - 1) ++cam_local[o.cam];
          stereo_local |= stereo;
          ++k;
        });
      }
      for (int c = 0; c < n_cam; ++c) cam_obs[t][c] = cam_local[c];
      stereo_seen[t] = stereo_local ? 1 : 0;
    });
    lap("fill");
    for (int t = 0; t < T; ++t) {
      for (int c = 0; c < n_cam; ++c) w.cam_obs[c] += cam_obs[t][c];
      w.any_stereo |= stereo_seen[t] != 0;
    }
  }
};

void finish(gpba_map& m, gpba_window& w) {
  gpba_problem& P = w.P;
  std::memset(&P, 0, sizeof(P));
  w.cam_intr = m.cam_intr; w.cam_Tbc = m.cam_Tbc;
  P.n_cam = m.n_cam; P.cam_intr = w.cam_intr.data(); P.cam_Tbc = w.cam_Tbc.data(); P.bf = m.bf;
  P.n_kf = (int32_t)w.kf_slot.size();
  w.kf_pose.resize((size_t)7 * P.n_kf); w.kf_vel.resize((size_t)6 * P.n_kf); w.kf_time.resize(P.n_kf); w.kf_id.resize(P.n_kf);
  for (int i = 0; i < P.n_kf; ++i) {
    const Kf& k = m.kfs[w.kf_slot[i]];
    std::memcpy(&w.kf_pose[7 * i], k.pose, sizeof(k.pose)); std::memcpy(&w.kf_vel[6 * i], k.vel, sizeof(k.vel));
    w.kf_time[i] = k.time; w.kf_id[i] = k.id;
  }
  P.kf_pose = w.kf_pose.data(); P.kf_vel = w.kf_vel.data(); P.kf_time = w.kf_time.data(); P.kf_fixed = w.kf_fixed.data();
  P.n_pt = (int32_t)w.pt_slot.size();
  w.pt_xyz.resize((size_t)3 * P.n_pt); w.pt_id.resize(P.n_pt);
  for (int i = 0; i < P.n_pt; ++i) { const Pt& p = m.pts[w.pt_slot[i]]; std::memcpy(&w.pt_xyz[3 * i], p.xyz, sizeof(p.xyz)); w.pt_id[i] = p.id; }
  P.pt_xyz = w.pt_xyz.data();
  P.n_rec = (int32_t)w.rec_kf2.size();
  P.rec_kf1 = w.rec_kf1.data(); P.rec_kf2 = w.rec_kf2.data(); P.rec_cam = w.rec_cam.data(); P.rec_t = w.rec_t.data();
  P.n_obs = (int64_t)w.obs_u.size();
  P.obs_u = w.obs_u.data(); P.obs_v = w.obs_v.data(); P.obs_ur = w.any_stereo ? w.obs_ur.data() : nullptr; P.obs_inv_sigma2 = w.obs_w.data();
  P.obs_rec = w.obs_rec.data(); P.obs_pt = w.obs_pt.data(); P.obs_flags = w.obs_flags.data();
  P.n_prior = (int32_t)w.prior_kf1.size(); P.prior_kf1 = w.prior_kf1.data(); P.prior_kf2 = w.prior_kf2.data();
  P.n_velp = (int32_t)w.velp_kf.size(); P.velp_kf = w.velp_kf.data();
  for (int i = 0; i < 6; ++i) P.qc[i] = m.qc[i];
  P.huber_mono = (double)(float)std::sqrt(5.991);    // thHuberMono (:957, :146)
  P.huber_stereo = (double)(float)std::sqrt(7.815);  // thHuberStereo
  P.linear_solver = GPBA_SOLVER_DENSE_CHOL;
}

}  // namespace

extern "C" {

int gpba_map_create(const gpba_map_config* cfg, gpba_map** out) {
  if (!cfg || !out || cfg->n_cam < 1 || cfg->n_cam > 63 || !cfg->cam_intr || !cfg->cam_Tbc) return fail("invalid map config");
  gpba_map* m = new gpba_map();
  m->n_cam = cfg->n_cam;
  m->cam_intr.assign(cfg->cam_intr, cfg->cam_intr + 4 * cfg->n_cam);
  m->cam_Tbc.assign(cfg->cam_Tbc, cfg->cam_Tbc + 7 * cfg->n_cam);
  m->bf = cfg->bf;
  for (int i = 0; i < 6; ++i) m->qc[i] = cfg->qc[i];
  *out = m;
  return GPBA_OK;
}
void gpba_map_destroy(gpba_map* m) { delete m; }

int gpba_map_add_keyframe(gpba_map* m, int64_t id, int64_t prev_id, const double pose[7], const double vel[6], double time, const double* cam_time) {
  if (!m || !pose || !vel || !cam_time) return fail("null argument");
  if (m->kf_of.count(id)) return fail("keyframe id already present");
  int prev = -1;
  if (prev_id >= 0) {
    auto it = m->kf_of.find(prev_id);
    if (it == m->kf_of.end() || m->kfs[it->second].bad) return fail("previous keyframe unknown");
    prev = it->second;
  }
  Kf k;
  k.id = id; k.prev = prev; k.time = time;
  std::memcpy(k.pose, pose, sizeof(k.pose)); std::memcpy(k.vel, vel, sizeof(k.vel));
  k.cam_time.assign(cam_time, cam_time + m->n_cam);
  const int slot = (int)m->kfs.size();
  m->kfs.push_back(std::move(k));
  m->kf_of.emplace(id, slot);
  if (prev >= 0) m->kfs[prev].next = slot;   // Tracking.cc:2217-2218
  ++m->n_kf_alive;
  return GPBA_OK;
}

int gpba_map_set_keyframe_state(gpba_map* m, int64_t id, const double pose[7], const double vel[6]) {
  if (!m) return fail("null argument");
  auto it = m->kf_of.find(id);
  if (it == m->kf_of.end()) return fail("keyframe unknown");
  Kf& k = m->kfs[it->second];
  if (pose) std::memcpy(k.pose, pose, sizeof(k.pose));
  if (vel) std::memcpy(k.vel, vel, sizeof(k.vel));
  return GPBA_OK;
}

int gpba_map_set_keyframe_bad(gpba_map* m, int64_t id) {
  if (!m) return fail("null argument");
  auto it = m->kf_of.find(id);
  if (it == m->kf_of.end()) return fail("keyframe unknown");
  Kf& k = m->kfs[it->second];
  if (k.bad) return GPBA_OK;
  if (k.prev >= 0 && k.next >= 0) {   // LocalMapping.cc:873-876
    m->kfs[k.next].prev = k.prev;
    m->kfs[k.prev].next = k.next;
    k.next = -1; k.prev = -1;
  }
  const int self = it->second;
  for (const auto& c : k.conn) {                                        // EraseConnection on every neighbour (:663-666, :763-777)
    Kf& o = m->kfs[c.first];
    if (o.conn.erase(self)) m->update_best_covisibles(o);
  }
  k.conn.clear(); k.ordered.clear();
  for (int s : k.obs) if (m->obs[s].alive) m->kill_obs(s, true, false);   // k.obs is being walked: compact afterwards
  m->compact(k);
  k.bad = true;
  --m->n_kf_alive;
  return GPBA_OK;
}

int gpba_map_add_point(gpba_map* m, int64_t id, const double xyz[3]) {
  if (!m || !xyz) return fail("null argument");
  if (m->pt_of.count(id)) return fail("point id already present");
  Pt p;
  p.id = id;
  std::memcpy(p.xyz, xyz, sizeof(p.xyz));
  m->pt_of.emplace(id, (int)m->pts.size());
  m->pts.push_back(std::move(p));
  ++m->n_pt_alive;
  return GPBA_OK;
}
int gpba_map_set_point(gpba_map* m, int64_t id, const double xyz[3]) {
  if (!m || !xyz) return fail("null argument");
  auto it = m->pt_of.find(id);
  if (it == m->pt_of.end()) return fail("point unknown");
  std::memcpy(m->pts[it->second].xyz, xyz, 3 * sizeof(double));
  return GPBA_OK;
}
int gpba_map_set_point_bad(gpba_map* m, int64_t id) {
  if (!m) return fail("null argument");
  auto it = m->pt_of.find(id);
  if (it == m->pt_of.end()) return fail("point unknown");
  Pt& p = m->pts[it->second];
  if (p.bad) return GPBA_OK;
  for (const PObs& o : p.obs) m->kill_obs(o.slot, false);
  p.obs.clear();
  p.bad = true;
  --m->n_pt_alive;
  return GPBA_OK;
}

int gpba_map_add_observation(gpba_map* m, int64_t kf, int32_t cam, int64_t pt, double u, double v, double ur, double w, int32_t close_flag) {
  if (!m) return fail("null argument");
  auto ik = m->kf_of.find(kf);
  auto ip = m->pt_of.find(pt);
  if (ik == m->kf_of.end() || ip == m->pt_of.end()) return fail("keyframe or point unknown");
  if (cam < 0 || cam >= m->n_cam) return fail("camera index out of range");
  if (m->kfs[ik->second].bad || m->pts[ip->second].bad) return fail("keyframe or point is bad");
  Pt& p = m->pts[ip->second];
  size_t pos;
  int s = m->find_obs(p, kf, cam, &pos);
  if (s < 0) {
    if (!m->free_obs.empty()) { s = m->free_obs.back(); m->free_obs.pop_back(); }
    else { s = (int)m->obs.size(); m->obs.emplace_back(); }
    p.obs.insert(p.obs.begin() + pos, PObs());
    m->kfs[ik->second].obs.push_back(s);
    ++m->n_obs_alive;
    Obs& o = m->obs[s];
    o.kf = ik->second; o.pt = ip->second; o.cam = cam; o.alive = true;
  }
  PObs& q = p.obs[pos];
  q.slot = s; q.kf = ik->second; q.cam = cam; q.close_flag = close_flag ? 1 : 0; q.u = u; q.v = v; q.ur = ur; q.w = w;
  return GPBA_OK;
}

int gpba_map_add_observations(gpba_map* m, int64_t n, const int64_t* kf, const int32_t* cam, const int64_t* pt, const double* u,
                              const double* v, const double* ur, const double* w, const uint8_t* close_flag, int64_t* n_done) {
  if (n_done) *n_done = 0;
  if (!m || n < 0 || (n > 0 && (!kf || !cam || !pt || !u || !v || !w))) return fail("null argument");
  for (int64_t i = 0; i < n; ++i) {
    const int rc = gpba_map_add_observation(m, kf[i], cam[i], pt[i], u[i], v[i], ur ? ur[i] : -1.0, w[i], close_flag ? close_flag[i] : 0);
    if (rc != GPBA_OK) return rc;
    if (n_done) *n_done = i + 1;
  }
  return GPBA_OK;
}

int gpba_map_erase_observation(gpba_map* m, int64_t kf, int32_t cam, int64_t pt) {
  if (!m) return fail("null argument");
  auto ip = m->pt_of.find(pt);
  if (ip == m->pt_of.end()) return fail("point unknown");
  const int s = m->find_obs(m->pts[ip->second], kf, cam, nullptr);
  if (s < 0) return fail("observation unknown");
  m->kill_obs(s, true);
  return GPBA_OK;
}

int gpba_map_update_connections(gpba_map* m, int64_t kf_id) {
  if (!m) return fail("null argument");
  auto it = m->kf_of.find(kf_id);
  if (it == m->kf_of.end() || m->kfs[it->second].bad) return fail("keyframe unknown");
  const int self = it->second;
  std::unordered_map<int, int> counter;                                 // KFcounter (:457-487)
  for (int s : m->kfs[self].obs) {
    const Obs& o = m->obs[s];
    if (!o.alive) continue;
    const Pt& p = m->pts[o.pt];
    if (p.bad) continue;
    int last = -1;
    for (const PObs& q : p.obs) {                                       // one count per observing keyframe
      if (q.kf == last) continue;
      last = q.kf;
      if (q.kf == self || m->kfs[q.kf].bad) continue;
      ++counter[q.kf];
    }
  }
  if (counter.empty()) return GPBA_OK;                                  // :490-491
  std::vector<std::pair<int64_t, int> > byid;                           // the reference walks a pointer-keyed map: by id here
  for (const auto& c : counter) byid.push_back({m->kfs[c.first].id, c.first});
  std::sort(byid.begin(), byid.end());
  const int th = 15;
  int nmax = 0, kmax = -1;
  std::vector<std::pair<std::pair<int, int64_t>, int> > pairs;
  for (const auto& e : byid) {
    const int w = counter[e.second];
    if (w > nmax) { nmax = w; kmax = e.second; }
    if (w >= th) { pairs.push_back({{w, e.first}, e.second}); m->add_connection(e.second, self, w); }
  }
  if (pairs.empty()) { pairs.push_back({{nmax, m->kfs[kmax].id}, kmax}); m->add_connection(kmax, self, nmax); }
  std::sort(pairs.begin(), pairs.end());
  Kf& k = m->kfs[self];
  k.conn = std::move(counter);                                          // ALL counted keyframes (:537)
  k.ordered.clear();
  for (size_t i = pairs.size(); i-- > 0;) k.ordered.push_back({pairs[i].second, pairs[i].first.first});
  return GPBA_OK;
}

int gpba_map_covisibles(const gpba_map* m, int64_t kf_id, int64_t* ids, int32_t* weights, int32_t capacity, int32_t* n_out) {
  if (!m || !n_out) return fail("null argument");
  auto it = m->kf_of.find(kf_id);
  if (it == m->kf_of.end()) return fail("keyframe unknown");
  const Kf& k = m->kfs[it->second];
  *n_out = (int32_t)k.ordered.size();
  for (int32_t i = 0; i < capacity && i < *n_out; ++i) {
    if (ids) ids[i] = m->kfs[k.ordered[i].first].id;
    if (weights) weights[i] = k.ordered[i].second;
  }
  return GPBA_OK;
}

int gpba_map_stats(const gpba_map* m, int64_t out[3]) {
  if (!m || !out) return fail("null argument");
  out[0] = m->n_kf_alive; out[1] = m->n_pt_alive; out[2] = m->n_obs_alive;
  return GPBA_OK;
}

int gpba_map_local_window(gpba_map* m, int64_t kf_id, int32_t large, const int64_t* covisible, int32_t n_cov, gpba_window** out) {
  if (!m || !out) return fail("null argument");
  auto it0 = m->kf_of.find(kf_id);
  if (it0 == m->kf_of.end() || m->kfs[it0->second].bad) return fail("keyframe unknown");
  const int64_t stamp = ++m->stamp;
  static const bool timing = std::getenv("GPBA_MAP_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "  [gpba_map] %-12s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  };
  const int maxOpt = large ? 25 : 10;                                   // :718-725
  const int Nd = (int)std::min<int64_t>(m->n_kf_alive - 2, maxOpt);
  std::vector<int> opt, vis, fixed, lpts;
  opt.push_back(it0->second);
  m->kfs[opt[0]].local_stamp = stamp;
  for (int i = 1; i < Nd; ++i) {                                        // :735-746
    const int p = m->kfs[opt.back()].prev;
    if (p < 0) break;
    opt.push_back(p);
    m->kfs[p].local_stamp = stamp;
  }
  auto add_points_of = [&](int kf) {                                    // :752-767, :797-810
    const Kf& k = m->kfs[kf];
    for (int s : k.obs) {
      const Obs& o = m->obs[s];
      if (!o.alive) continue;
      Pt& p = m->pts[o.pt];
      if (p.bad || p.stamp == stamp) continue;
      p.stamp = stamp;
      lpts.push_back(o.pt);
    }
  };
  for (int kf : opt) add_points_of(kf);
  {                                                                     // :770-782
    const int last = opt.back();
    if (m->kfs[last].prev >= 0) {
      fixed.push_back(m->kfs[last].prev);
      m->kfs[fixed.back()].fixed_stamp = stamp;
    } else {
      m->kfs[last].local_stamp = -1;
      m->kfs[last].fixed_stamp = stamp;
      fixed.push_back(last);
      opt.pop_back();
    }
  }
  std::vector<int64_t> own_cov;
  if (n_cov < 0) {                                                      // pKF->GetVectorCovisibleKeyFrames() from the mirror
    for (const auto& c : m->kfs[it0->second].ordered) own_cov.push_back(m->kfs[c.first].id);
    covisible = own_cov.data(); n_cov = (int32_t)own_cov.size();
  }
  for (int i = 0; i < n_cov; ++i) {                                     // :784-812 (maxCovKF = 0: one keyframe at most)
    if (!vis.empty()) break;
    auto it = m->kf_of.find(covisible[i]);
    if (it == m->kf_of.end()) continue;
    Kf& k = m->kfs[it->second];
    if (k.local_stamp == stamp || k.fixed_stamp == stamp) continue;
    k.local_stamp = stamp;
    if (k.bad) continue;
    vis.push_back(it->second);
    add_points_of(it->second);
  }
  lap("select");
  const size_t maxFixKF = 50;                                           // :815-836
  // sequential by definition (a point adds at most the first keyframe nobody has marked yet), and a pointer chase: every
  // point's observation list is its own heap block.  The loads are prefetched two hops ahead (point header, then list).
  const size_t nl = lpts.size();
  for (size_t li = 0; li < nl; ++li) {
    if (li + 16 < nl) __builtin_prefetch(&m->pts[lpts[li + 16]]);
    if (li + 8 < nl) __builtin_prefetch(m->pts[lpts[li + 8]].obs.data());
    const int ps = lpts[li];
    const Pt& p = m->pts[ps];
    int last_kf = -1;
    for (const PObs& o : p.obs) {
      const int kf = o.kf;
      if (kf == last_kf) continue;
      last_kf = kf;
      Kf& k = m->kfs[kf];
      if (k.local_stamp != stamp && k.fixed_stamp != stamp) {
        k.fixed_stamp = stamp;
        if (!k.bad) { fixed.push_back(kf); break; }
      }
    }
    if (fixed.size() >= maxFixKF) break;
  }

  lap("fixed scan");
  gpba_window* w = new_window();
  w->local = true; w->large = large != 0; w->iterations = 10;          // opt_it1 (:1221)
  w->cam_obs.assign(m->n_cam, 0);
  // vertices in ascending id = Hessian order (sparse_optimizer.cpp:166-190)
  struct Member { int slot; int role; };
  std::vector<Member> mem;
  for (int k : opt) mem.push_back({k, 0});
  for (int k : vis) mem.push_back({k, 1});
  for (int k : fixed) mem.push_back({k, 2});
  std::sort(mem.begin(), mem.end(), [&](const Member& a, const Member& b) { return m->kfs[a.slot].id < m->kfs[b.slot].id; });
  for (size_t i = 0; i < mem.size(); ++i) {
    m->kfs[mem[i].slot].widx = (int)i;
    w->kf_slot.push_back(mem[i].slot); w->kf_role.push_back(mem[i].role); w->kf_fixed.push_back(mem[i].role == 2 ? 1 : 0);
  }
  for (int k : opt) w->velp_kf.push_back(m->kfs[k].widx);               // EdgeVelocity on the temporal keyframes (:869-872)
  for (int i = (int)opt.size() - 1; i > 0; --i) {                       // EdgeGaussianPrior (:897-910)
    w->prior_kf1.push_back(m->kfs[opt[i]].widx);
    w->prior_kf2.push_back(m->kfs[opt[i - 1]].widx);
  }
  auto in_graph = [&](int kf) {
    const Kf& k = m->kfs[kf];
    return (!k.bad && (k.local_stamp == stamp || k.fixed_stamp == stamp)) ? k.widx : -1;
  };
  w->pt_slot.assign(lpts.begin(), lpts.end());
  Emitter(*m, *w).run(lpts, mem.size(), in_graph);
  lap("emit");
  finish(*m, *w);
  lap("finish");
  w->P.huber_prior = 0.0;                                               // kernel commented out (:906-908)
  w->P.lambda_init = large ? 1e-2 : 1.0;                                // :846, :852
  *out = w;
  return GPBA_OK;
}

int gpba_map_global_window(gpba_map* m, int64_t init_kf_id, gpba_window** out) {
  if (!m || !out) return fail("null argument");
  gpba_window* w = new_window();
  w->local = false; w->iterations = 10;                                 // LoopClosing.cc:1221
  w->cam_obs.assign(m->n_cam, 0);
  std::vector<int> order;
  for (size_t i = 0; i < m->kfs.size(); ++i) if (!m->kfs[i].bad) order.push_back((int)i);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return m->kfs[a].id < m->kfs[b].id; });
  for (Kf& k : m->kfs) k.widx = -1;
  for (size_t i = 0; i < order.size(); ++i) {
    Kf& k = m->kfs[order[i]];
    k.widx = (int)i;
    w->kf_slot.push_back(order[i]); w->kf_role.push_back(k.id == init_kf_id ? 2 : 0); w->kf_fixed.push_back(k.id == init_kf_id ? 1 : 0);
  }
  for (int s : order) {                                                 // :101-135
    const Kf& k = m->kfs[s];
    w->velp_kf.push_back(k.widx);
    if (k.prev >= 0 && !m->kfs[k.prev].bad) { w->prior_kf1.push_back(m->kfs[k.prev].widx); w->prior_kf2.push_back(k.widx); }
  }
  std::vector<int> porder;
  for (size_t i = 0; i < m->pts.size(); ++i) if (!m->pts[i].bad) porder.push_back((int)i);
  std::sort(porder.begin(), porder.end(), [&](int a, int b) { return m->pts[a].id < m->pts[b].id; });
  auto in_graph = [&](int kf) { return m->kfs[kf].bad ? -1 : m->kfs[kf].widx; };
  for (int ps : porder)
    if (!m->pts[ps].obs.empty()) w->pt_slot.push_back(ps);              // nEdges == 0: removeVertex (:305-309); bad keyframes hold no observations
  Emitter(*m, *w).run(w->pt_slot, order.size(), in_graph);
  finish(*m, *w);
  w->P.huber_prior = 21.026;                                            // :128-130
  w->P.lambda_init = 1e-5;                                              // :75
  *out = w;
  return GPBA_OK;
}

void gpba_window_destroy(gpba_window* w) {
  if (!w) return;
  {
    std::lock_guard<std::mutex> lock(g_spare_mutex);
    if (g_spare.size() < 2) { g_spare.push_back(w); return; }
  }
  delete w;
}
const gpba_problem* gpba_window_problem(const gpba_window* w) { return w ? &w->P : nullptr; }
int32_t gpba_window_iterations(const gpba_window* w) { return w ? w->iterations : 0; }

int gpba_window_ids(const gpba_window* w, int64_t* kf_id, int32_t* kf_role, int64_t* pt_id, int64_t* obs_kf, int32_t* obs_cam, int64_t* obs_pt) {
  if (!w) return fail("null argument");
  if (kf_id) std::copy(w->kf_id.begin(), w->kf_id.end(), kf_id);
  if (kf_role) std::copy(w->kf_role.begin(), w->kf_role.end(), kf_role);
  if (pt_id) std::copy(w->pt_id.begin(), w->pt_id.end(), pt_id);
  for (size_t i = 0; i < w->obs_rec.size(); ++i) {
    const int r = w->obs_rec[i];
    if (obs_kf) obs_kf[i] = w->kf_id[w->rec_kf2[r]];
    if (obs_cam) obs_cam[i] = w->rec_cam[r];
    if (obs_pt) obs_pt[i] = w->pt_id[w->obs_pt[i]];
  }
  return GPBA_OK;
}
int gpba_window_cam_obs(const gpba_window* w, int32_t* cam_obs) {
  if (!w || !cam_obs) return fail("null argument");
  std::copy(w->cam_obs.begin(), w->cam_obs.end(), cam_obs);
  return GPBA_OK;
}

int gpba_window_apply(gpba_map* m, const gpba_window* w, const double* kf_pose, const double* kf_vel, const double* pt_xyz,
                      const uint8_t* flags, float err, float err_end, int32_t* applied, int64_t* n_erased, int64_t* erased_obs) {
  if (!m || !w) return fail("null argument");
  if (n_erased) *n_erased = 0;
  if (applied) *applied = 0;
  if (w->local && !w->large && (2 * err < err_end || std::isnan(err) || std::isnan(err_end))) return GPBA_OK;   // :1354-1358
  int64_t ne = 0;
  if (flags) {
    for (size_t i = 0; i < w->obs_slot.size(); ++i) {
      if (!flags[i]) continue;
      const int s = w->obs_slot[i];
      // the slot may have been erased (or erased and re-used) by the map since the window was built: only erase the
      // observation the window saw
      const Obs& o = m->obs[s];
      const int r = w->obs_rec[i];
      if (!o.alive || m->kfs[o.kf].id != w->kf_id[w->rec_kf2[r]] || o.cam != w->rec_cam[r] || m->pts[o.pt].id != w->pt_id[w->obs_pt[i]]) continue;
      if (m->pts[o.pt].bad) continue;                                   // "if (pMP->isBad()) continue" (:1276 ...)
      m->kill_obs(s, true);
      if (erased_obs) erased_obs[ne] = (int64_t)i;
      ++ne;
    }
  }
  if (kf_pose) {
    for (size_t i = 0; i < w->kf_slot.size(); ++i) {
      if (w->kf_fixed[i]) continue;
      Kf& k = m->kfs[w->kf_slot[i]];
      double q[4];
      double nq = 0;
      for (int c = 0; c < 4; ++c) { q[c] = f32(kf_pose[7 * i + c]); nq += q[c] * q[c]; }
      nq = std::sqrt(nq);
      for (int c = 0; c < 4; ++c) k.pose[c] = q[c] / nq;
      for (int c = 4; c < 7; ++c) k.pose[c] = f32(kf_pose[7 * i + c]);
      if (kf_vel) for (int c = 0; c < 6; ++c) k.vel[c] = f32(kf_vel[6 * i + c]);
    }
  }
  if (pt_xyz) {
    for (size_t i = 0; i < w->pt_slot.size(); ++i) {
      Pt& p = m->pts[w->pt_slot[i]];
      for (int c = 0; c < 3; ++c) p.xyz[c] = f32(pt_xyz[3 * i + c]);
    }
  }
  if (applied) *applied = 1;
  if (n_erased) *n_erased = ne;
  return GPBA_OK;
}

// MultiKeyFrame::mTbc[c] = v->estimate().cast<float>() for the cameras that were calibrated (src/Optimizer.cc:1419-1428)
int gpba_map_apply_extrinsics(gpba_map* m, const gpba_window* w, const double* cam_Tbc, int32_t min_obs, int32_t* n_updated) {
  if (!m || !w || !cam_Tbc) return fail("null argument");
  int32_t n = 0;
  for (int c = 0; c < m->n_cam - 1; ++c) {          // the last camera is the synchronous reference camera: no VertexExtrinsic (:983)
    if (w->cam_obs[c] < min_obs) continue;           // "if (cam_obs[c] < extrin_thresh) continue" (:1421-1424)
    double q[4], nq = 0;
    for (int k = 0; k < 4; ++k) { q[k] = f32(cam_Tbc[7 * c + k]); nq += q[k] * q[k]; }
    nq = std::sqrt(nq);
    if (!(nq > 0)) return fail("degenerate extrinsic quaternion");
    for (int k = 0; k < 4; ++k) m->cam_Tbc[7 * c + k] = q[k] / nq;   // SE3f -> cast<double>() re-normalises (so3.hpp:480-487)
    for (int k = 4; k < 7; ++k) m->cam_Tbc[7 * c + k] = f32(cam_Tbc[7 * c + k]);
    ++n;
  }
  if (n_updated) *n_updated = n;
  return GPBA_OK;
}
int gpba_map_extrinsics(const gpba_map* m, double* cam_Tbc) {
  if (!m || !cam_Tbc) return fail("null argument");
  std::copy(m->cam_Tbc.begin(), m->cam_Tbc.end(), cam_Tbc);
  return GPBA_OK;
}

const char* gpba_map_last_error(void) { return g_map_err.c_str(); }

}  // extern "C"
