// dqn.cpp — implementation of the dqn::DQN mirror on top of the C-ABI (include/dqn_b200.h).
// Control flow, flags, RNG call order, log-line formats and error convention (CHECK/LOG(FATAL)
// abort) follow the reference's src/dqn.cpp; the arithmetic is in libdqn_b200.so.
#include "dqn.hpp"
#include "caffe_proto.hpp"

#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <regex>

#include "../../include/dqn_b200.h"
#include "shim/flags.hpp"
#include "shim/logging.hpp"

namespace dqn {

using namespace hfo;

struct DQN::ReplayGroup {
  std::mutex mu;                    // guards `members`
  std::vector<DQN *> members;
};
struct DQN::ShareGroup {
  std::mutex mu;
  std::vector<DQN *> members;
  int n_actor = 0, n_critic = 0;
};

// dqn.cpp:21-31
DEFINE_int32(seed, 0, "Seed the RNG. Default: time");
DEFINE_double(tau, .001, "Step size for soft updates.");
DEFINE_int32(soft_update_freq, 1, "Do SoftUpdateNet this frequently");
DEFINE_double(gamma, .99, "Discount factor of future rewards (0,1]");
DEFINE_int32(memory, 500000, "Capacity of replay memory");
DEFINE_int32(memory_threshold, 1000, "Number of transitions required to start learning");
DEFINE_int32(loss_display_iter, 1000, "Frequency of loss display");
DEFINE_int32(snapshot_freq, 10000, "Frequency (steps) snapshots");
DEFINE_bool(remove_old_snapshots, true, "Remove old snapshots when writing more recent ones.");
DEFINE_bool(snapshot_memory, true, "Snapshot the replay memory along with the network.");
DEFINE_double(beta, .5, "Mix between off-policy and on-policy updates.");
// run-time replacements for the reference's compile-time constants (dqn.hpp:19, dqn.cpp:425,:449)
DEFINE_int32(batch_size, kMinibatchSize, "Minibatch size of Update() (kMinibatchSize upstream)");
DEFINE_string(hidden, "1024,512,256,128", "Tower widths of both nets");
DEFINE_int32(device, 0, "CUDA device ordinal");
DEFINE_bool(host_sampling, false, "Draw minibatch indices on the host with std::mt19937 exactly like "
                                  "SampleTransitionsFromMemory (default: device Philox sampler)");
DEFINE_double(init_std, 0.01, "Std of the gaussian weight filler (dqn.cpp:352)");
DEFINE_bool(caffe_snapshots, true, "Write .caffemodel / .solverstate as Caffe protobufs (NetParameter / SolverState, what "
                                   "Solver::Snapshot writes upstream).  -nocaffe_snapshots writes this build's flat "
                                   "DQNBW001 / DQNBS001 files instead.  Reading detects either format");
DEFINE_bool(async_update, false, "Update() enqueues the update and books the loss of the PREVIOUS one (dqnb_update_async / "
                                 "dqnb_results): episodes, AddTransitions and logging overlap the GPU work.  The sampled "
                                 "memories, weights and iteration counts are those of the blocking loop; only the "
                                 "smoothed-loss log lags by one update");

#define DQNB_OK(call)                                                         \
  do {                                                                        \
    if ((call) != 0) LOG(FATAL) << #call << " failed: " << dqnb_last_error(); \
  } while (0)

static std::vector<int> parse_hidden(const std::string &s) {
  std::vector<int> out;
  size_t i = 0;
  while (i < s.size()) {
    size_t j = s.find(',', i);
    if (j == std::string::npos) j = s.size();
    out.push_back(std::atoi(s.substr(i, j - i).c_str()));
    i = j + 1;
  }
  return out;
}

// ---- free functions ---------------------------------------------------------------------------
// Offset of a param of a given action, -1 if there is none (dqn.cpp:162-178).
static int GetParamOffset(const action_t action, const int arg_num = 0) {
  if (arg_num < 0 || arg_num > 1) return -1;
  switch (action) {
    case DASH: return arg_num;
    case TURN: return arg_num == 0 ? 2 : -1;
    case TACKLE: return arg_num == 0 ? 3 : -1;
    case KICK: return 4 + arg_num;
    default: LOG(FATAL) << "Unrecognized action: " << action;
  }
  return -1;
}

static Action ActionFromChoice(action_t a, const ActorOutput &o) {
  Action action;
  action.action = a;
  const int o1 = GetParamOffset(a, 0);
  CHECK_GE(o1, 0);
  action.arg1 = o[kActionSize + o1];
  const int o2 = GetParamOffset(a, 1);
  action.arg2 = o2 < 0 ? 0 : o[kActionSize + o2];
  return action;
}

Action GetAction(const ActorOutput &actor_output) {  // dqn.cpp:196-208
  ActorOutput copy(actor_output);
  copy[TACKLE] = -99999;  // tackle is never chosen
  const action_t best = (action_t)std::distance(copy.begin(), std::max_element(copy.begin(), copy.begin() + kActionSize));
  return ActionFromChoice(best, actor_output);
}

Action DQN::SampleAction(const ActorOutput &actor_output) {  // dqn.cpp:180-194
  const float dash = std::max(0., actor_output[DASH] + 1.0), turn = std::max(0., actor_output[TURN] + 1.0);
  const float tackle = 0, kick = std::max(0., actor_output[KICK] + 1.0);
  std::discrete_distribution<int> dist{dash, turn, tackle, kick};
  return ActionFromChoice((action_t)dist(random_engine), actor_output);
}

std::string PrintActorOutput(const ActorOutput &o) {  // dqn.cpp:210-216
  return "Dash(" + std::to_string(o[4]) + ", " + std::to_string(o[5]) + ")=" + std::to_string(o[0]) + ", Turn(" +
         std::to_string(o[6]) + ")=" + std::to_string(o[1]) + ", Tackle(" + std::to_string(o[7]) + ")=" +
         std::to_string(o[2]) + ", Kick(" + std::to_string(o[8]) + ", " + std::to_string(o[9]) + ")=" + std::to_string(o[3]);
}

caffe::NetParameter CreateActorNet(int state_size) {
  caffe::NetParameter np;
  np.set_name("Actor"); np.set_force_backward(true);
  np.state_size = state_size; np.critic = false; np.hidden = parse_hidden(FLAGS_hidden);
  return np;
}
caffe::NetParameter CreateCriticNet(int state_size) {
  caffe::NetParameter np;
  np.set_name("Critic"); np.set_force_backward(true);
  np.state_size = state_size; np.critic = true; np.hidden = parse_hidden(FLAGS_hidden);
  return np;
}

static bool is_regular_file(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }

std::vector<std::string> FilesMatchingRegexp(const std::string &regexp) {  // dqn.cpp:559-580
  std::string dir = ".", stem = regexp;
  const size_t slash = regexp.find_last_of('/');
  if (slash != std::string::npos) { dir = regexp.substr(0, slash); stem = regexp.substr(slash + 1); if (dir.empty()) dir = "/"; }
  std::vector<std::string> out;
  const std::regex re(stem);
  if (DIR *d = opendir(dir.c_str())) {
    while (dirent *e = readdir(d)) {
      const std::string name = e->d_name, full = (slash == std::string::npos ? name : dir + "/" + name);
      if (is_regular_file(full) && std::regex_match(name, re)) out.push_back(full);
    }
    closedir(d);
  }
  return out;
}
static int ParseIterFromSnapshot(const std::string &s) {
  const size_t a = s.find_last_of('_'), b = s.find_last_of('.');
  return std::stoi(s.substr(a + 1, b - a - 1));
}
void RemoveFilesMatchingRegexp(const std::string &regexp) {
  for (const std::string &f : FilesMatchingRegexp(regexp)) { LOG(INFO) << "Removing " << f; std::remove(f.c_str()); }
}
void RemoveSnapshots(const std::string &regexp, int min_iter) {
  for (const std::string &f : FilesMatchingRegexp(regexp))
    if (ParseIterFromSnapshot(f) < min_iter) { LOG(INFO) << "Removing " << f; std::remove(f.c_str()); }
}
static int FindGreatestIter(const std::string &regexp) {
  int mx = -1;
  for (const std::string &f : FilesMatchingRegexp(regexp)) mx = std::max(mx, ParseIterFromSnapshot(f));
  return mx;
}
void FindLatestSnapshot(const std::string &prefix, std::string &actor_snapshot, std::string &critic_snapshot,
                        std::string &memory_snapshot) {  // dqn.cpp:122-144
  const int a = FindGreatestIter(prefix + "_actor_iter_[0-9]+\\.solverstate");
  const int c = FindGreatestIter(prefix + "_critic_iter_[0-9]+\\.solverstate");
  const int m = FindGreatestIter(prefix + "_iter_[0-9]+\\.replaymemory");
  if (a > 0) actor_snapshot = prefix + "_actor_iter_" + std::to_string(a) + ".solverstate";
  if (c > 0) critic_snapshot = prefix + "_critic_iter_" + std::to_string(c) + ".solverstate";
  if (m > 0) memory_snapshot = prefix + "_iter_" + std::to_string(m) + ".replaymemory";
}
int FindHiScore(const std::string &prefix) {  // dqn.cpp:146-158
  int best = std::numeric_limits<int>::lowest();
  for (const std::string &f : FilesMatchingRegexp(prefix + "_HiScore[-]?[0-9]+_iter_[0-9]+\\.caffemodel")) {
    const size_t a = f.find("_HiScore"), b = f.find("_iter_");
    best = std::max(best, std::stoi(f.substr(a + 8, b - a - 8)));
  }
  return best;
}

// ---- construction -------------------------------------------------------------------------------
DQN::DQN(caffe::SolverParameter &actor_solver_param, caffe::SolverParameter &critic_solver_param,
         std::string save_path, int state_size, int tid)
    : actor_solver_param_(actor_solver_param), critic_solver_param_(critic_solver_param),
      replay_memory_capacity_(FLAGS_memory), gamma_(FLAGS_gamma), h_(nullptr), random_engine(),
      smoothed_critic_loss_(0), smoothed_actor_loss_(0), last_snapshot_iter_(0), save_path_(save_path),
      state_size_(state_size), batch_size_(FLAGS_batch_size), tid_(tid), unum_(0), actor_iter_cache_(0),
      critic_iter_cache_(0), iters_dirty_(true), last_update_(0.f, 0.f) {
  unsigned seed = FLAGS_seed;
  if (FLAGS_seed <= 0) {  // dqn.cpp:474-481
    seed = (unsigned)std::chrono::system_clock::now().time_since_epoch().count();
    LOG(INFO) << "Seeding RNG to time (seed = " << seed << ")";
  } else {
    LOG(INFO) << "Seeding RNG with seed = " << FLAGS_seed;
  }
  random_engine.seed(seed);
  // Initialize (dqn.cpp:622-662): both solvers/nets + the two target clones live behind one handle.
  dqnb_config c;
  dqnb_default_config(&c);
  c.device = FLAGS_device; c.state_size = state_size; c.batch = batch_size_;
  // the tower comes with the solver parameters (CreateActorNet / a user's .prototxt, dqn_main.cpp:232-246);
  // both nets share one width list in this build
  const std::vector<int> hidden = actor_solver_param.net_param_.hidden.empty() ? parse_hidden(FLAGS_hidden)
                                                                                : actor_solver_param.net_param_.hidden;
  if (!critic_solver_param.net_param_.hidden.empty())
    CHECK(critic_solver_param.net_param_.hidden == hidden) << "actor and critic towers of different widths are not implemented";
  if (actor_solver_param.net_param_.state_size > 0) CHECK_EQ(actor_solver_param.net_param_.state_size, state_size);
  if (critic_solver_param.net_param_.state_size > 0) CHECK_EQ(critic_solver_param.net_param_.state_size, state_size);
  CHECK_LE((int)hidden.size(), DQNB_MAX_HIDDEN);
  hidden_ = hidden;
  c.n_hidden = (int)hidden.size();
  for (size_t i = 0; i < hidden.size(); ++i) c.hidden[i] = hidden[i];
  c.replay_capacity = replay_memory_capacity_;
  c.max_act_batch = std::max(batch_size_, 1);   // SelectActions accepts up to batch_size_ states (dqn.cpp:699)
  c.gamma = FLAGS_gamma; c.beta = FLAGS_beta; c.tau = (float)FLAGS_tau; c.soft_update_freq = FLAGS_soft_update_freq;
  CHECK(actor_solver_param.type() == "Adam" && critic_solver_param.type() == "Adam") << "only the Adam solver is implemented";
  c.actor_lr = actor_solver_param.base_lr(); c.critic_lr = critic_solver_param.base_lr();
  c.momentum = actor_solver_param.momentum(); c.momentum2 = actor_solver_param.momentum2();
  c.delta = actor_solver_param.delta(); c.clip_gradients = actor_solver_param.clip_gradients();
  c.seed = seed;
  DQNB_OK(dqnb_create(&c, &h_));
  // gaussian(0.01) weight fill + CloneNet x2 (dqn.cpp:350-352, :660-661)
  DQNB_OK(dqnb_init_params(h_, seed, (float)FLAGS_init_std));
}

DQN::~DQN() {
  if (replay_share_) {
    std::lock_guard<std::mutex> lock(replay_share_->mu);
    auto &m = replay_share_->members;
    m.erase(std::remove(m.begin(), m.end(), this), m.end());
  }
  if (share_) {          // leave the sharing group: the others must not write through into a dead handle
    std::lock_guard<std::mutex> lock(share_->mu);
    std::lock_guard<std::recursive_mutex> self(mu_);
    auto &m = share_->members;
    m.erase(std::remove(m.begin(), m.end(), this), m.end());
  }
  drain_pending();
  dqnb_destroy(h_);
}

// -async_update: the loss of the last enqueued update has not been looked at yet (dqn.cpp:906 checks every one)
void DQN::drain_pending() {
  std::lock_guard<std::recursive_mutex> self(mu_);

  if (pending_step_ <= 0) return;
  float loss = 0.f, avg_q = 0.f;
  DQNB_OK(dqnb_results(h_, pending_step_, 1, &loss, &avg_q));
  pending_step_ = 0;
  CHECK(std::isfinite(loss)) << "Critic loss not finite!";
}

void DQN::refresh_iters() const {
  std::lock_guard<std::recursive_mutex> self(mu_);

  if (!iters_dirty_) return;
  int32_t a = 0, c = 0;
  DQNB_OK(dqnb_iters(h_, &a, &c));
  actor_iter_cache_ = a; critic_iter_cache_ = c; iters_dirty_ = false;
}
int DQN::critic_iter() const { refresh_iters(); return critic_iter_cache_; }
int DQN::actor_iter() const { refresh_iters(); return actor_iter_cache_; }
int DQN::memory_size() const { std::lock_guard<std::recursive_mutex> self(mu_); return dqnb_memory_size(h_); }
void DQN::ClearReplayMemory() {
  for (DQN *m : replay_targets()) {
    std::lock_guard<std::recursive_mutex> lock(m->mu_);
    DQNB_OK(dqnb_clear_memory(m->h_));
  }
}

void DQN::Benchmark(int iterations) {
  std::lock_guard<std::recursive_mutex> self(mu_);
  // dqn.cpp:487-498
  LOG(INFO) << "*** Benchmark begins ***";
  drain_pending();
  float ms = 0.f;
  DQNB_OK(dqnb_benchmark(h_, iterations, &ms));
  iters_dirty_ = true;
  LOG(INFO) << "Average Update: " << ms / iterations << " ms.";
  LOG(INFO) << "*** Benchmark ends ***";
}

// ---- acting -------------------------------------------------------------------------------------
ActorOutput DQN::GetRandomActorOutput() {  // dqn.cpp:664-682: ten draws in this order
  ActorOutput o;
  auto U = [&](float lo, float hi) { return std::uniform_real_distribution<float>(lo, hi)(random_engine); };
  for (int i = 0; i < kActionSize; ++i) o[i] = U(-1.0, 1.0);
  o[kActionSize + 0] = U(-100.0, 100.0);  // dash power
  o[kActionSize + 1] = U(-180.0, 180.0);  // dash angle
  o[kActionSize + 2] = U(-180.0, 180.0);  // turn angle
  o[kActionSize + 3] = U(-180.0, 180.0);  // tackle angle
  o[kActionSize + 4] = U(0.0, 100.0);     // kick power
  o[kActionSize + 5] = U(-180.0, 180.0);  // kick angle
  return o;
}

ActorOutput DQN::SelectAction(const InputStates &last_states, const double epsilon) {
  return SelectActions(std::vector<InputStates>{{last_states}}, epsilon)[0];
}

std::vector<ActorOutput> DQN::SelectActions(const std::vector<InputStates> &states_batch, const double epsilon) {
  std::lock_guard<std::recursive_mutex> self(mu_);

  CHECK(epsilon >= 0.0 && epsilon <= 1.0);
  CHECK_LE((int)states_batch.size(), batch_size_);
  std::vector<ActorOutput> out(states_batch.size());
  if (std::uniform_real_distribution<double>(0.0, 1.0)(random_engine) < epsilon) {  // one coin flip per batch
    for (auto &o : out) o = GetRandomActorOutput();
    return out;
  }
  // SelectActionGreedily (dqn.cpp:734-766)
  const int n = (int)states_batch.size();
  std::vector<float> flat((size_t)n * state_size_), res((size_t)n * 10);
  for (int i = 0; i < n; ++i) {
    const StateDataSp &s = states_batch[i][kStateInputCount - 1];
    CHECK_EQ((int)s->size(), state_size_);
    std::copy(s->begin(), s->end(), flat.begin() + (size_t)i * state_size_);
  }
  DQNB_OK(dqnb_select_actions(h_, n, flat.data(), res.data()));
  for (int i = 0; i < n; ++i) std::copy(res.begin() + i * 10, res.begin() + (i + 1) * 10, out[i].begin());
  return out;
}

float DQN::EvaluateAction(const InputStates &input_states, const ActorOutput &action) {
  std::lock_guard<std::recursive_mutex> self(mu_);
  // dqn.cpp:688-693
  float q = 0.f;
  DQNB_OK(dqnb_evaluate(h_, 1, input_states[kStateInputCount - 1]->data(), action.data(), &q));
  return q;
}

// ---- replay memory ------------------------------------------------------------------------------
std::vector<DQN *> DQN::replay_targets() {
  std::shared_ptr<ReplayGroup> grp;
  { std::lock_guard<std::recursive_mutex> self(mu_); grp = replay_share_; }
  if (!grp) return {this};
  std::lock_guard<std::mutex> lock(grp->mu);
  return grp->members;
}
// rows go into every ring of the replay group, one member at a time (no two object mutexes held together)
void DQN::add_rows(int n, const float *s, const float *a, const float *r, const float *mc, const float *sn, const uint8_t *term) {
  for (DQN *m : replay_targets()) {
    std::lock_guard<std::recursive_mutex> lock(m->mu_);
    DQNB_OK(dqnb_add_transitions(m->h_, n, s, a, r, mc, sn, term));
  }
}

void DQN::AddTransition(const Transition &t) {   // dqn.cpp:768-773 (evicts when size == capacity)
  const auto &next = std::get<4>(t);
  for (DQN *m : replay_targets()) {
    std::lock_guard<std::recursive_mutex> lock(m->mu_);
    DQNB_OK(dqnb_add_transition(m->h_, std::get<0>(t)[kStateInputCount - 1]->data(), std::get<1>(t).data(), std::get<2>(t),
                                std::get<3>(t), next ? (*next)->data() : nullptr, next ? 0 : 1));
  }
}

void DQN::AddTransitions(const std::vector<Transition> &ts) {
  const int n = (int)ts.size();
  if (n == 0) return;
  std::vector<float> s((size_t)n * state_size_), sn((size_t)n * state_size_, 0.f), a((size_t)n * 10), r(n), mc(n);
  std::vector<uint8_t> term(n);
  for (int i = 0; i < n; ++i) {
    const Transition &t = ts[i];
    std::copy(std::get<0>(t)[kStateInputCount - 1]->begin(), std::get<0>(t)[kStateInputCount - 1]->end(), s.begin() + (size_t)i * state_size_);
    std::copy(std::get<1>(t).begin(), std::get<1>(t).end(), a.begin() + (size_t)i * 10);
    r[i] = std::get<2>(t); mc[i] = std::get<3>(t);
    const auto &next = std::get<4>(t);
    term[i] = next ? 0 : 1;   // dqn.cpp:878: terminal <=> no next state
    if (next) std::copy((*next)->begin(), (*next)->end(), sn.begin() + (size_t)i * state_size_);
  }
  add_rows(n, s.data(), a.data(), r.data(), mc.data(), sn.data(), term.data());
}

void DQN::LabelTransitions(std::vector<Transition> &transitions) {  // dqn.cpp:783-797
  CHECK_GT(transitions.size(), 0u) << "Need at least one transition to label.";
  Transition &last = transitions[transitions.size() - 1];
  std::get<3>(last) = std::get<2>(last);
  for (int i = (int)transitions.size() - 2; i >= 0; --i) {
    Transition &t = transitions[i];
    const float reward = std::get<2>(t), target = std::get<3>(transitions[i + 1]);
    std::get<3>(t) = reward + gamma_ * target;   // double gamma_, narrowed on store
  }
}

// ---- learning -----------------------------------------------------------------------------------
std::vector<int> DQN::SampleTransitionsFromMemory(int n) {  // dqn.cpp:501-509
  std::vector<int> idx(n);
  const int size = memory_size();
  for (int i = 0; i < n; ++i) idx[i] = std::uniform_int_distribution<int>(0, size - 1)(random_engine);
  return idx;
}

std::pair<float, float> DQN::UpdateActorCritic() {
  std::lock_guard<std::recursive_mutex> self(mu_);

  float loss = 0.f, avg_q = 0.f;
  if (FLAGS_async_update && !FLAGS_host_sampling) {
    long long step = 0;
    DQNB_OK(dqnb_update_async(h_, 1, (int64_t *)&step));
    if (!iters_dirty_) { actor_iter_cache_ += 1; critic_iter_cache_ += 1; }   // Solver::Step ++iter_, set_iter(iter+1)
    if (pending_step_ > 0) {                     // the update enqueued by the previous call has had a whole env step
      DQNB_OK(dqnb_results(h_, pending_step_, 1, &loss, &avg_q));
      CHECK(std::isfinite(loss)) << "Critic loss not finite!";   // dqn.cpp:906
    }
    pending_step_ = step;
    return std::make_pair(loss, avg_q);
  }
  if (FLAGS_host_sampling) {
    const std::vector<int> idx = SampleTransitionsFromMemory(batch_size_);
    DQNB_OK(dqnb_update_with_indices(h_, idx.data(), &loss, &avg_q));
  } else {
    DQNB_OK(dqnb_update(h_, 1, &loss, &avg_q));
  }
  if (!iters_dirty_) { actor_iter_cache_ += 1; critic_iter_cache_ += 1; }   // no device read per update
  CHECK(std::isfinite(loss)) << "Critic loss not finite!";   // dqn.cpp:906
  return std::make_pair(loss, avg_q);
}

void DQN::Update() {  // dqn.cpp:799-826
  if (memory_size() < FLAGS_memory_threshold) return;
  std::shared_ptr<ShareGroup> grp;
  { std::lock_guard<std::recursive_mutex> self(mu_); grp = share_; }   // thread 0 may be attaching us to its group right now
  if (grp) {             // multi-agent sharing: one member updates at a time, then its shared layers reach the others
    // lock order: the group, then this object, then one teammate at a time (a teammate that is acting holds only its
    // own mutex, for the length of one call)
    std::lock_guard<std::mutex> lock(grp->mu);
    std::lock_guard<std::recursive_mutex> self(mu_);
    last_update_ = UpdateActorCritic();
    drain_pending();
    for (DQN *m : grp->members)
      if (m != this) {
        std::lock_guard<std::recursive_mutex> mate(m->mu_);
        m->drain_pending();
        DQNB_OK(dqnb_copy_shared_layers(m->h_, h_, grp->n_actor, grp->n_critic));
      }
  } else
  last_update_ = UpdateActorCritic();
  if (critic_iter() % FLAGS_loss_display_iter == 0) {
    LOG(INFO) << "[Agent" << tid_ << "] Critic Iteration " << critic_iter() << ", loss = " << smoothed_critic_loss_;
    smoothed_critic_loss_ = 0;
  }
  smoothed_critic_loss_ += last_update_.first / float(FLAGS_loss_display_iter);
  if (actor_iter() % FLAGS_loss_display_iter == 0) {
    LOG(INFO) << "[Agent" << tid_ << "] Actor Iteration " << actor_iter() << ", avg_q_value = " << smoothed_actor_loss_;
    smoothed_actor_loss_ = 0;
  }
  smoothed_actor_loss_ += last_update_.second / float(FLAGS_loss_display_iter);
  if (critic_iter() >= last_snapshot_iter_ + FLAGS_snapshot_freq || actor_iter() >= last_snapshot_iter_ + FLAGS_snapshot_freq) {
    Snapshot();
    last_snapshot_iter_ = max_iter();
  }
}

// ---- snapshots ----------------------------------------------------------------------------------
// .caffemodel / .solverstate are Caffe protobufs upstream (NetParameter / SolverState).  Both are read
// (caffe_proto.cpp: a reference checkpoint loads as it is); written on request (-caffe_snapshots), the default
// being flat little-endian files:
//   caffemodel : "DQNBW001" int64 n, float w[n]            (Caffe learnable_params order)
//   solverstate: "DQNBS001" int32 iter, int64 n, float m[n], float v[n], then the caffemodel payload
static void write_blob(const std::string &f, const char *magic, int32_t iter, const std::vector<const std::vector<float> *> &arrs) {
  std::ofstream o(f, std::ios::binary);
  CHECK(o.good()) << "cannot write " << f;
  o.write(magic, 8);
  o.write((const char *)&iter, 4);
  const int64_t n = (int64_t)arrs[0]->size();
  o.write((const char *)&n, 8);
  for (auto *a : arrs) o.write((const char *)a->data(), sizeof(float) * a->size());
}
static void read_blob(const std::string &f, const char *magic, int32_t *iter, std::vector<std::vector<float>> &arrs, int count, int64_t expect_n) {
  std::ifstream in(f, std::ios::binary);
  CHECK(in.good()) << "Invalid file: " << f;
  char m[8];
  in.read(m, 8);
  CHECK(std::memcmp(m, magic, 8) == 0) << f << " is not a " << magic << " file";
  in.read((char *)iter, 4);
  int64_t n = 0;
  in.read((char *)&n, 8);
  CHECK(in.good() && n == expect_n) << f << " holds " << n << " parameters, this net has " << expect_n;
  arrs.assign(count, std::vector<float>((size_t)n));
  for (auto &a : arrs) in.read((char *)a.data(), sizeof(float) * n);
  CHECK(in.good()) << "truncated file " << f;
}

static std::string read_file(const std::string &f) {
  std::ifstream in(f, std::ios::binary);
  CHECK(in.good()) << "Invalid file: " << f;
  std::stringstream ss;
  ss << in.rdbuf();
  return ss.str();
}
static void write_file(const std::string &f, const std::string &bytes) {
  std::ofstream o(f, std::ios::binary);
  CHECK(o.good()) << "cannot write " << f;
  o.write(bytes.data(), (std::streamsize)bytes.size());
}

void DQN::snapshot_net(int net, const std::string &base) const {
  std::lock_guard<std::recursive_mutex> self(mu_);

  const int64_t n = dqnb_param_count(h_, net);
  std::vector<float> w((size_t)n), m((size_t)n), v((size_t)n);
  int32_t iter = 0;
  DQNB_OK(dqnb_get_params(h_, net, w.data()));
  DQNB_OK(dqnb_get_opt_state(h_, net, m.data(), v.data(), &iter));
  if (FLAGS_caffe_snapshots) {   // Solver::SnapshotToBinaryProto + SnapshotSolverStateToBinaryProto
    const bool critic = net == DQNB_CRITIC;
    write_file(base + ".caffemodel", caffe_proto::EncodeNet(caffe_proto::NetFromFlat(critic ? "Critic" : "Actor", state_size_, hidden_, critic, w.data())));
    caffe_proto::SolverState st;
    st.iter = iter;
    st.learned_net = base + ".caffemodel";
    st.history = caffe_proto::HistoryFromFlat(caffe_proto::ParamLayers(state_size_, hidden_, critic), m.data(), v.data());
    write_file(base + ".solverstate", caffe_proto::EncodeSolverState(st));
    return;
  }
  write_blob(base + ".caffemodel", "DQNBW001", iter, {&w});
  write_blob(base + ".solverstate", "DQNBS001", iter, {&m, &v, &w});
}

// Net::CopyTrainedLayersFrom (dqn.cpp:529,:537): this build's flat file, or a Caffe NetParameter whose layers are
// matched by name (layers the net does not have are ignored, layers the file does not have keep their weights)
void DQN::load_weights(int net, const std::string &f) {
  std::lock_guard<std::recursive_mutex> self(mu_);

  const std::string bytes = read_file(f);
  const int64_t n = dqnb_param_count(h_, net);
  std::vector<float> w((size_t)n);
  if (bytes.size() >= 8 && std::memcmp(bytes.data(), "DQNBW001", 8) == 0) {
    std::vector<std::vector<float>> a; int32_t it;
    read_blob(f, "DQNBW001", &it, a, 1, n);
    CHECK_EQ((int64_t)a[0].size(), n);
    w = a[0];
  } else {
    caffe_proto::Net proto;
    CHECK(caffe_proto::DecodeNet(bytes, &proto)) << f << " is neither a DQNBW001 file nor a Caffe NetParameter";
    DQNB_OK(dqnb_get_params(h_, net, w.data()));
    std::string err;
    const bool critic = net == DQNB_CRITIC;
    const int copied = caffe_proto::FlatFromNet(proto, state_size_, hidden_, critic, w.data(), &err);
    CHECK_GE(copied, 0) << f << ": " << err;
    LOG(INFO) << "Copied " << copied << " of " << caffe_proto::ParamLayers(state_size_, hidden_, critic).size()
              << " parametrised layers from " << f << " (net \"" << proto.name << "\", " << proto.layers.size() << " layers)";
  }
  DQNB_OK(dqnb_set_params(h_, net, w.data()));
  DQNB_OK(dqnb_set_params(h_, net + 2, w.data()));   // CloneNet: targets are re-cloned, not checkpointed (dqn.cpp:530,:538,:546,:555)
}

void DQN::Snapshot() { Snapshot(save_path_, FLAGS_remove_old_snapshots, FLAGS_snapshot_memory); }

void DQN::Snapshot(const std::string &prefix, bool remove_old, bool snapshot_memory) {
  std::lock_guard<std::recursive_mutex> self(mu_);
  // dqn.cpp:586-620
  drain_pending();
  const int ai = actor_iter(), ci = critic_iter();
  snapshot_net(DQNB_ACTOR, prefix + "_actor_iter_" + std::to_string(ai));
  snapshot_net(DQNB_CRITIC, prefix + "_critic_iter_" + std::to_string(ci));
  if (snapshot_memory) {
    const std::string mem = prefix + "_iter_" + std::to_string(max_iter()) + ".replaymemory";
    LOG(INFO) << "Snapshotting memory to " << mem;
    SnapshotReplayMemory(mem);
    CHECK(is_regular_file(mem));
  }
  if (remove_old) {
    RemoveSnapshots(prefix + "_actor_iter_[0-9]+\\.(caffemodel|solverstate)", ai - 1);
    RemoveSnapshots(prefix + "_critic_iter_[0-9]+\\.(caffemodel|solverstate)", ci - 1);
    RemoveSnapshots(prefix + "_iter_[0-9]+\\.replaymemory", ci - 1);
  }
  LOG(INFO) << "Snapshotting Finished!";
}

void DQN::LoadActorWeights(const std::string &f) { load_weights(DQNB_ACTOR, f); }    // dqn.cpp:525-531
void DQN::LoadCriticWeights(const std::string &f) { load_weights(DQNB_CRITIC, f); }  // dqn.cpp:533-539

// Solver::Restore (dqn.cpp:541-557): iteration + Adam history, weights from the model the state points to
void DQN::restore_solver(int net, const std::string &f) {
  std::lock_guard<std::recursive_mutex> self(mu_);

  const std::string bytes = read_file(f);
  const int64_t n = dqnb_param_count(h_, net);
  if (bytes.size() >= 8 && std::memcmp(bytes.data(), "DQNBS001", 8) == 0) {
    std::vector<std::vector<float>> a; int32_t it;
    read_blob(f, "DQNBS001", &it, a, 3, n);
    CHECK_EQ((int64_t)a[0].size(), n);
    DQNB_OK(dqnb_set_params(h_, net, a[2].data()));
    DQNB_OK(dqnb_set_params(h_, net + 2, a[2].data()));   // targets are re-cloned, not checkpointed (dqn.cpp:546,:555)
    DQNB_OK(dqnb_set_opt_state(h_, net, a[0].data(), a[1].data(), it));
    return;
  }
  caffe_proto::SolverState st;
  CHECK(caffe_proto::DecodeSolverState(bytes, &st)) << f << " is neither a DQNBS001 file nor a Caffe SolverState";
  if (!st.learned_net.empty()) {
    std::string model = st.learned_net;
    if (!is_regular_file(model)) {   // snapshots moved to another directory: look beside the solver state
      const size_t sl = f.find_last_of('/'), ml = model.find_last_of('/');
      model = (sl == std::string::npos ? std::string() : f.substr(0, sl + 1)) + (ml == std::string::npos ? model : model.substr(ml + 1));
    }
    load_weights(net, model);
  }
  std::vector<float> m((size_t)n), v((size_t)n);
  std::string err;
  CHECK(caffe_proto::FlatFromHistory(st.history, caffe_proto::ParamLayers(state_size_, hidden_, net == DQNB_CRITIC), m.data(), v.data(), &err))
      << f << ": " << err;
  DQNB_OK(dqnb_set_opt_state(h_, net, m.data(), v.data(), st.iter));
}
void DQN::RestoreActorSolver(const std::string &f) {
  LOG(INFO) << "Actor solver state resuming from " << f;
  restore_solver(DQNB_ACTOR, f); iters_dirty_ = true; last_snapshot_iter_ = max_iter();
}
void DQN::RestoreCriticSolver(const std::string &f) {
  LOG(INFO) << "Critic solver state resuming from " << f;
  restore_solver(DQNB_CRITIC, f); iters_dirty_ = true; last_snapshot_iter_ = max_iter();
}

// Replay-memory file: the reference's gzip layout (dqn.cpp:1152-1173, kStateInputCount == 1):
//   int32 n, then per transition  state[S] f32 | ActorOutput 10 f32 | reward f32 | on_policy_target f32 | terminal u8
// with the next state implicit = the following record's state unless terminal (dqn.cpp:1218-1220).
void DQN::SnapshotReplayMemory(const std::string &filename) {
  std::lock_guard<std::recursive_mutex> self(mu_);

  gzFile out = gzopen(filename.c_str(), "wb");
  CHECK(out != nullptr) << "cannot write " << filename;
  const int n = memory_size();
  gzwrite(out, &n, sizeof(int));
  const int chunk = 8192;
  std::vector<float> s((size_t)chunk * state_size_), a((size_t)chunk * 10), r(chunk), mc(chunk);
  std::vector<uint8_t> t(chunk);
  int episodes = 0;
  for (int first = 0; first < n; first += chunk) {
    const int m = std::min(chunk, n - first);
    DQNB_OK(dqnb_get_transitions(h_, first, m, s.data(), a.data(), r.data(), mc.data(), nullptr, t.data()));
    for (int i = 0; i < m; ++i) {
      gzwrite(out, s.data() + (size_t)i * state_size_, sizeof(float) * state_size_);
      gzwrite(out, a.data() + (size_t)i * 10, sizeof(float) * 10);
      gzwrite(out, &r[i], sizeof(float));
      gzwrite(out, &mc[i], sizeof(float));
      const bool terminal = t[i] != 0;
      gzwrite(out, &terminal, sizeof(bool));
      if (terminal) episodes++;
    }
  }
  gzclose(out);
  LOG(INFO) << "Saved memory of size " << n << " with " << episodes << " episodes";
}

void DQN::LoadReplayMemory(const std::string &filename) {
  CHECK(is_regular_file(filename)) << "Invalid file: " << filename;
  LOG(INFO) << "Loading replay memory from " << filename;
  ClearReplayMemory();
  gzFile in = gzopen(filename.c_str(), "rb");
  CHECK(in != nullptr);
  int n = 0;
  CHECK_EQ(gzread(in, &n, sizeof(int)), (int)sizeof(int));
  std::vector<float> s((size_t)n * state_size_), sn((size_t)n * state_size_, 0.f), a((size_t)n * 10), r(n), mc(n);
  std::vector<uint8_t> t(n);
  int episodes = 0;
  for (int i = 0; i < n; ++i) {
    CHECK_EQ(gzread(in, s.data() + (size_t)i * state_size_, sizeof(float) * state_size_), (int)(sizeof(float) * state_size_));
    gzread(in, a.data() + (size_t)i * 10, sizeof(float) * 10);
    gzread(in, &r[i], sizeof(float));
    gzread(in, &mc[i], sizeof(float));
    bool terminal = true;
    gzread(in, &terminal, sizeof(bool));
    t[i] = terminal ? 1 : 0;
    if (terminal) episodes++;
  }
  gzclose(in);
  // the last record of a file cut mid-episode has no successor: upstream leaves it without a next
  // state (boost::none), i.e. terminal for the update (dqn.cpp:1217-1220)
  for (int i = 0; i < n; ++i) {
    if (!t[i] && i + 1 < n) std::copy(s.begin() + (size_t)(i + 1) * state_size_, s.begin() + (size_t)(i + 2) * state_size_, sn.begin() + (size_t)i * state_size_);
    else t[i] = 1;
  }
  // upstream resizes the deque to n regardless of capacity; the ring holds at most capacity-1 via AddTransitions
  const int chunk = std::max(1, std::min(n, replay_memory_capacity_ / 2));
  for (int first = 0; first < n; first += chunk) {
    const int m = std::min(chunk, n - first);
    add_rows(m, s.data() + (size_t)first * state_size_, a.data() + (size_t)first * 10, r.data() + first,
             mc.data() + first, sn.data() + (size_t)first * state_size_, t.data() + first);
  }
  LOG(INFO) << "replay_mem_size = " << memory_size() << " with " << episodes << " episodes";
}

void DQN::ShareLayer(caffe::Layer<float> &, caffe::Layer<float> &) {
  LOG(FATAL) << "ShareLayer (dqn.cpp:1037-1046) takes Caffe layer objects, which do not exist in this build; use ShareParameters";
}
// dqn.cpp:1048-1079.  Upstream the slave's blobs ShareData() with the owner's: one copy of the first n layers (and of the
// same layers of the target nets) that every agent's solver updates in turn, each with its own Adam history.  Here every
// DQN owns a handle with flat device buffers, so a group keeps the copies identical by writing through: `other` takes
// the owner's layers now, and after each member's UpdateActorCritic its shared layers are copied to the other members
// (DQN::Update, under the group's mutex: upstream's agent threads race on the shared blobs, this serialises them).
void DQN::ShareParameters(DQN &other, int num_actor_layers_to_share, int num_critic_layers_to_share) {
  CHECK(&other != this);
  CHECK_GE(num_actor_layers_to_share, 0);
  CHECK_GE(num_critic_layers_to_share, 0);
  CHECK_LE(num_actor_layers_to_share, (int)hidden_.size() + 2) << "the actor has " << hidden_.size() + 2 << " layers with parameters";
  CHECK_LE(num_critic_layers_to_share, (int)hidden_.size() + 1) << "the critic has " << hidden_.size() + 1 << " layers with parameters";
  if (!share_) {
    share_ = std::make_shared<ShareGroup>();
    share_->members.push_back(this);
    share_->n_actor = num_actor_layers_to_share;
    share_->n_critic = num_critic_layers_to_share;
  }
  CHECK(share_->n_actor == num_actor_layers_to_share && share_->n_critic == num_critic_layers_to_share)
      << "one sharing group shares one set of layers (dqn_main.cpp:311 passes the same flags for every teammate)";
  std::lock_guard<std::mutex> lock(share_->mu);
  static const char *kActorNames[] = {"action_layer", "actionpara_layer"};
  for (int i = 0; i < num_actor_layers_to_share; ++i)
    LOG(INFO) << "Sharing Actor Layer " << (i < (int)hidden_.size() ? "ip" + std::to_string(i + 1) + "_layer" : std::string(kActorNames[i - hidden_.size()]));
  for (int i = 0; i < num_critic_layers_to_share; ++i)
    LOG(INFO) << "Sharing Critic Layer " << (i < (int)hidden_.size() ? "ip" + std::to_string(i + 1) + "_layer" : std::string("q_values_layer"));
  std::lock_guard<std::recursive_mutex> self(mu_);
  std::lock_guard<std::recursive_mutex> mate(other.mu_);     // the teammate's thread may already be playing
  drain_pending();
  other.drain_pending();
  DQNB_OK(dqnb_copy_shared_layers(other.h_, h_, num_actor_layers_to_share, num_critic_layers_to_share));
  other.share_ = share_;
  share_->members.push_back(&other);
}
// dqn.cpp:1081-1083: `other.replay_memory_ = replay_memory_` - replay_memory_ is a shared_ptr (dqn.hpp:187), so from here
// on the two DQNs use ONE deque (dqn_main.cpp:146-149, :359-363 wrap its users in a global mutex).  `other`'s ring takes
// the owner's contents now and joins the owner's replay group: every later append / clear / load reaches every member.
void DQN::ShareReplayMemory(DQN &other) {
  CHECK(&other != this);
  CHECK_EQ(state_size_, other.state_size_);
  CHECK_EQ(replay_memory_capacity_, other.replay_memory_capacity_) << "a shared replay memory has one capacity";
  std::shared_ptr<ReplayGroup> grp;
  {
    std::lock_guard<std::recursive_mutex> self(mu_);
    if (!replay_share_) {
      replay_share_ = std::make_shared<ReplayGroup>();
      replay_share_->members.push_back(this);
    }
    grp = replay_share_;
  }
  std::lock_guard<std::mutex> glock(grp->mu);                 // no member list changes / group-wide appends meanwhile
  const int n = memory_size(), S = state_size_, chunk = 16384;
  std::vector<float> s((size_t)chunk * S), sn((size_t)chunk * S), a((size_t)chunk * (kActionSize + kActionParamSize)), r(chunk), mc(chunk);
  std::vector<uint8_t> term(chunk);
  {
    std::lock_guard<std::recursive_mutex> mate(other.mu_);
    DQNB_OK(dqnb_clear_memory(other.h_));
  }
  for (int first = 0; first < n; first += chunk) {
    const int m = std::min(chunk, n - first);
    {
      std::lock_guard<std::recursive_mutex> self(mu_);
      DQNB_OK(dqnb_get_transitions(h_, first, m, s.data(), a.data(), r.data(), mc.data(), sn.data(), term.data()));
    }
    std::lock_guard<std::recursive_mutex> mate(other.mu_);
    DQNB_OK(dqnb_add_transitions(other.h_, m, s.data(), a.data(), r.data(), mc.data(), sn.data(), term.data()));
  }
  std::lock_guard<std::recursive_mutex> mate(other.mu_);
  other.replay_share_ = grp;
  grp->members.push_back(&other);
}

}  // namespace dqn
