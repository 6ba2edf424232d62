// hill_climb.h -- the hill-climbing enumerator and accept loop as one state machine, usable on the host
// (lock-step particle rounds, particles.cu) and on the device (the whole match in one launch, score.cu).
//
// FailedRoundsLimitedPoseEnumerator<Distorsion1DPoseEnumerator> + the accept loop of
// PoseEnumerationScanMatcher, one instance per matcher (hill_climbing_scan_matcher.h:10-126,
// pose_enumeration_scan_matcher.h:48-65); frame rotation is always 0 upstream (quirk Q5)
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SG_HC_HD __host__ __device__
#else
#define SG_HC_HD
#endif

struct HillClimb {
  double bx, by, bt, best;     // best pose so far and its probability
  double rbx, rby, rbt;        // base of the current round
  double tr, rot;
  unsigned failed_rounds = 0, action = 0;
  bool base_set = false, round_failed = true, done = false;
  int64_t tested = 1;

  // candidates the reference would test next, up to the end of the current round
  SG_HC_HD int next_round(unsigned max_failed_rounds, double out[6][3]) {
    int k = 0;
    unsigned fr = failed_rounds, act = action;
    bool bs = base_set, rf = round_failed;
    double t_tr = tr, t_rot = rot, x0 = rbx, y0 = rby, t0 = rbt;
    while (fr < max_failed_rounds && k < 6) {
      if (!(act < 6)) {
        if (k > 0) break;  // the next round depends on this round's accepts: stop here
        if (rf) { t_tr *= 0.5; t_rot *= 0.5; ++fr; }
        act = 0; bs = false; rf = true;
      }
      if (!bs) { x0 = bx; y0 = by; t0 = bt; bs = true; }
      double x = x0, y = y0, t = t0;
      const double dir = act % 2 ? -1 : 1;
      switch (act % 3) {
        case 0: x += 1.0 * dir * t_tr; y += 0.0 * dir * t_tr; break;
        case 1: x += -0.0 * dir * t_tr; y += 1.0 * dir * t_tr; break;
        case 2: t += dir * t_rot; break;
      }
      ++act;
      out[k][0] = x; out[k][1] = y; out[k][2] = t;
      ++k;
    }
    return k;
  }
  // replay the same steps with the scores known
  SG_HC_HD void apply(unsigned max_failed_rounds, const double cand[6][3], const double *scores, int k) {
    for (int j = 0; j < k; ++j) {
      if (!(action < 6)) {
        if (round_failed) { tr *= 0.5; rot *= 0.5; ++failed_rounds; }
        action = 0; base_set = false; round_failed = true;
      }
      if (!base_set) { rbx = bx; rby = by; rbt = bt; base_set = true; }
      ++action;
      ++tested;
      const bool ok = best < scores[j];
      round_failed &= !ok;
      if (ok) { best = scores[j]; bx = cand[j][0]; by = cand[j][1]; bt = cand[j][2]; }
    }
    done = !(failed_rounds < max_failed_rounds);
  }
};
