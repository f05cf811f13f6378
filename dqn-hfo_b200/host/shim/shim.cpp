// shim.cpp — implementations behind the gflags/glog/HFO stand-ins.
#include <cmath>
#include <cstring>
#include <iostream>

#include "HFO.hpp"
#include "caffe/caffe.hpp"
#include "flags.hpp"
#include "gflags/gflags.h"
#include "glog/logging.h"
#include "logging.hpp"

// knobs of the in-process environment (the real ones are arguments of the HFO server: hfo_game.cpp:13)
DEFINE_int32(frames_per_trial, 500, "Episode length cap of the in-process HFO stand-in (--frames-per-trial upstream)");
DEFINE_int32(env_seed, 1, "Seed of the in-process HFO stand-in");

namespace fLI { int FLAGS_logbuflevel = 0; }

namespace gflags {
static std::string g_usage, g_version;
void SetUsageMessage(const std::string &usage) { g_usage = usage; }
void SetVersionString(const std::string &version) { g_version = version; }
const char *ProgramUsage() { return g_usage.c_str(); }
}  // namespace gflags

namespace caffe {
static Caffe::Brew g_mode = Caffe::GPU;
void Caffe::set_mode(Brew mode) {
  if (mode == CPU && g_mode != CPU) LOG(WARNING) << "Caffe::set_mode(CPU) ignored: the update path of this build runs on the GPU only";
  g_mode = mode;
}
Caffe::Brew Caffe::mode() { return g_mode; }
}  // namespace caffe

namespace shim {
int g_vlog_level = 0;

// value of an int32 flag that another translation unit DEFINEd (e.g. the caller's -offense_agents), or dflt
int int_flag_or(const char *name, int dflt) {
  auto it = flag_registry().find(name);
  return (it != flag_registry().end() && it->second.type == 'i') ? *static_cast<int32_t *>(it->second.ptr) : dflt;
}

int ParseCommandLineFlags(int *argc, char ***argv, bool remove_flags) {
  int kept = 1, used = 0;
  char **av = *argv;
  for (int i = 1; i < *argc; ++i) {
    std::string a = av[i];
    if (a.size() < 2 || a[0] != '-') { av[kept++] = av[i]; continue; }
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string name = a, value;
    bool has_value = false;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) { name = a.substr(0, eq); value = a.substr(eq + 1); has_value = true; }
    if (name == "v") { g_vlog_level = has_value ? std::atoi(value.c_str()) : 1; ++used; continue; }
    auto it = flag_registry().find(name);
    if (it == flag_registry().end() && name.rfind("no", 0) == 0) {
      auto it2 = flag_registry().find(name.substr(2));
      if (it2 != flag_registry().end() && it2->second.type == 'b') { set_flag(name.substr(2), "false"); ++used; continue; }
    }
    if (it == flag_registry().end()) { std::cerr << "ERROR: unknown command line flag '" << name << "'\n"; std::exit(1); }
    if (!has_value) {
      if (it->second.type == 'b') value = "true";
      else if (i + 1 < *argc) value = av[++i];
      else { std::cerr << "ERROR: flag '" << name << "' needs a value\n"; std::exit(1); }
    }
    set_flag(name, value);
    ++used;
  }
  if (remove_flags) *argc = kept;
  return used;
}
}  // namespace shim

namespace hfo {

std::string ActionToString(action_t a) {
  static const char *n[] = {"Dash", "Turn", "Tackle", "Kick", "KickTo", "MoveTo", "DribbleTo", "Intercept",
                            "Move", "Shoot", "Pass", "Dribble", "Catch", "No-op", "Quit"};
  return (a >= DASH && a <= QUIT) ? n[a] : "Unknown";
}
std::string StatusToString(status_t s) {
  static const char *n[] = {"InGame", "Goal", "CapturedByDefense", "OutOfBounds", "OutOfTime", "ServerDown"};
  return (s >= IN_GAME && s <= SERVER_DOWN) ? n[s] : "Unknown";
}

// Field: x in [0,1] towards the goal at (1,0), y in [-0.7,0.7].  Proximities are 1 - distance/2,
// angles relative to the agent's heading, as in the low-level feature set's conventions
// (features in [-1,1]).  Only the indices HFOGameState::update reads are meaningful
// (hfo_game.cpp:130-152: 12 kickable, 13-15 goal sin/cos/proximity, 51-53 ball sin/cos/proximity,
// 54-55 ball velocity validity/magnitude); the rest is deterministic filler derived from the pose.
HFOEnvironment::HFOEnvironment() : rng_(1) { feat_.assign(num_features_, 0.f); reset_episode(); }

void HFOEnvironment::configure(int num_features, int frames_per_trial, unsigned seed) {
  num_features_ = num_features; frames_per_trial_ = frames_per_trial; rng_.seed(seed);
  configured_ = true;
  feat_.assign(num_features_, 0.f);
  reset_episode();
}
// The real server tells the agent how many players are on the pitch; here the caller's own flags do
// (dqn_main.cpp:52-58: NumStateFeatures = 50 + 9 * players, hfo_game.hpp:14-16).
void HFOEnvironment::connectToServer(feature_set_t, std::string, int, std::string, std::string, bool, std::string) {
  int players = 0;
  for (const char *f : {"offense_agents", "offense_npcs", "offense_dummies", "defense_agents", "defense_npcs", "defense_dummies",
                        "defense_chasers"})
    players += shim::int_flag_or(f, 0);
  if (players > 0 && !configured_) configure(50 + 9 * players, FLAGS_frames_per_trial, (unsigned)FLAGS_env_seed);
}

void HFOEnvironment::reset_episode() {
  std::uniform_real_distribution<float> ux(0.05f, 0.3f), uy(-0.3f, 0.3f), ua(-3.14159f, 3.14159f);
  px_ = ux(rng_); py_ = uy(rng_); heading_ = ua(rng_);
  bx_ = px_ + 0.15f + 0.2f * ux(rng_); by_ = uy(rng_); bvx_ = bvy_ = 0.f;
  frame_ = 0; fresh_ = true; pending_ = NOOP;
  on_ball_ = Player();
  fill_features();
}

static inline float wrap(float a) { while (a > 3.14159265f) a -= 6.2831853f; while (a < -3.14159265f) a += 6.2831853f; return a; }

void HFOEnvironment::fill_features() {
  const float dbx = bx_ - px_, dby = by_ - py_, dgx = 1.f - px_, dgy = 0.f - py_;
  const float bd = std::sqrt(dbx * dbx + dby * dby), gd = std::sqrt(dgx * dgx + dgy * dgy);
  const float ba = wrap(std::atan2(dby, dbx) - heading_), ga = wrap(std::atan2(dgy, dgx) - heading_);
  for (int i = 0; i < num_features_; ++i) feat_[i] = std::sin(0.7f * i + 3.f * px_ - 2.f * py_ + heading_) * 0.5f;
  feat_[12] = bd < 0.04f ? 1.f : -1.f;
  feat_[13] = std::sin(ga); feat_[14] = std::cos(ga); feat_[15] = 1.f - std::min(gd / 2.f, 1.f) * 2.f * 0.5f;
  feat_[51] = std::sin(ba); feat_[52] = std::cos(ba); feat_[53] = 1.f - std::min(bd / 2.f, 1.f) * 2.f * 0.5f;
  feat_[54] = 1.f; feat_[55] = std::min(std::sqrt(bvx_ * bvx_ + bvy_ * bvy_) * 4.f - 1.f, 1.f);
}

const std::vector<float> &HFOEnvironment::getState() { return feat_; }
void HFOEnvironment::act(action_t action, float arg1, float arg2) { pending_ = action; arg1_ = arg1; arg2_ = arg2; }
Player HFOEnvironment::playerOnBall() { return on_ball_; }

status_t HFOEnvironment::step() {
  if (fresh_) {}  // first step of an episode: the pose drawn by reset_episode() is used as is
  if (frame_ < 0) reset_episode();   // the previous step ended an episode
  const float deg = 3.14159265f / 180.f;
  switch (pending_) {
    case DASH: {
      const float pw = std::max(-100.f, std::min(100.f, arg1_)) / 100.f * 0.03f;
      const float dir = heading_ + arg2_ * deg;
      px_ += pw * std::cos(dir); py_ += pw * std::sin(dir);
      break;
    }
    case TURN: heading_ = wrap(heading_ + arg1_ * deg); break;
    case KICK: {
      const float dbx = bx_ - px_, dby = by_ - py_;
      if (std::sqrt(dbx * dbx + dby * dby) < 0.04f) {
        const float pw = std::max(0.f, std::min(100.f, arg1_)) / 100.f * 0.08f;
        const float dir = heading_ + arg2_ * deg;
        bvx_ = pw * std::cos(dir); bvy_ = pw * std::sin(dir);
        on_ball_.side = LEFT; on_ball_.unum = unum_;
      }
      break;
    }
    default: break;
  }
  bx_ += bvx_; by_ += bvy_; bvx_ *= 0.94f; bvy_ *= 0.94f;
  {
    const float dbx = bx_ - px_, dby = by_ - py_;
    if (std::sqrt(dbx * dbx + dby * dby) < 0.04f) { on_ball_.side = LEFT; on_ball_.unum = unum_; }
  }
  pending_ = NOOP;
  fresh_ = false;
  ++frame_;
  status_t st = IN_GAME;
  if (bx_ >= 1.f && std::fabs(by_) < 0.2f) st = GOAL;
  else if (bx_ < 0.f || bx_ > 1.f || std::fabs(by_) > 0.7f || px_ < -0.1f || px_ > 1.1f || std::fabs(py_) > 0.8f) st = OUT_OF_BOUNDS;
  else if (frame_ >= frames_per_trial_) st = OUT_OF_TIME;
  fill_features();
  if (st != IN_GAME) frame_ = -1;    // terminal observation stays readable; next step() starts a new episode
  return st;
}

}  // namespace hfo
