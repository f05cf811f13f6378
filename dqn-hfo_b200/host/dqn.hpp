// dqn.hpp — host-side mirror of the reference's learner class (reference src/dqn.hpp:16-244): same
// namespace, type aliases, member names and argument meaning, so dqn_main-style callers compile
// against it unchanged.  Where the reference owns four caffe::Net and two caffe::Solver objects,
// this class owns one opaque dqnb_handle (include/dqn_b200.h): replay memory, weights, Adam state
// and target nets live in HBM and every FLOP runs in libdqn_b200.so.
//
// The reference's own src/dqn_main.cpp compiles against this header UNMODIFIED (host/Makefile target `dqn`,
// tests/test_host.py): the same includes resolve to the stand-ins under shim/ (<boost/optional.hpp> is
// std::optional, <caffe/caffe.hpp> the two protobuf messages of the interface as plain structs, glog / gflags
// look-alikes, <HFO.hpp> an in-process environment), because Boost, Caffe, glog, gflags and libhfo are not in
// this image.  kMinibatchSize stays the reference's compile-time default but the minibatch actually used
// is the run-time flag -batch_size (BASELINE configs need 1024/4096/8192).
#ifndef DQN_HPP_
#define DQN_HPP_

#include <HFO.hpp>
#include <caffe/caffe.hpp>
#include <boost/functional/hash.hpp>
#include <boost/optional.hpp>
#include <array>
#include <deque>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "hfo_game.hpp"

struct dqnb_handle_s;

namespace dqn {

constexpr auto kStateInputCount = 1;       // dqn.hpp:18
constexpr auto kMinibatchSize = 32;        // dqn.hpp:19 (default of -batch_size)
constexpr auto kActionSize = 4;            // dqn.hpp:20
constexpr auto kActionParamSize = 6;       // dqn.hpp:21

// Layer / blob names of the two nets (dqn.hpp:38-51); they appear in the .prototxt files
constexpr auto state_input_layer_name = "state_input_layer";
constexpr auto action_input_layer_name = "action_input_layer";
constexpr auto action_params_input_layer_name = "action_params_input_layer";
constexpr auto target_input_layer_name = "target_input_layer";
constexpr auto q_values_layer_name = "q_values_layer";
constexpr auto states_blob_name = "states";
constexpr auto actions_blob_name = "actions";
constexpr auto action_params_blob_name = "action_params";
constexpr auto targets_blob_name = "target";
constexpr auto q_values_blob_name = "q_values";
constexpr auto loss_blob_name = "loss";

using ActorOutput = std::array<float, kActionSize + kActionParamSize>;
using StateData = std::vector<float>;
using StateDataSp = std::shared_ptr<StateData>;
using InputStates = std::array<StateDataSp, kStateInputCount>;
using Transition = std::tuple<InputStates, ActorOutput, float, float, boost::optional<StateDataSp>>;
using SolverSp = std::shared_ptr<caffe::Solver<float>>;
using NetSp = boost::shared_ptr<caffe::Net<float>>;

class DQN {
 public:
  DQN(caffe::SolverParameter &actor_solver_param, caffe::SolverParameter &critic_solver_param,
      std::string save_path, int state_size, int tid);
  ~DQN();

  // Benchmark the speed of updates (dqn.cpp:487-498)
  void Benchmark(int iterations = 1000);

  // Loading methods (dqn.cpp:525-557, :1180-1226)
  void RestoreActorSolver(const std::string &actor_solver);
  void RestoreCriticSolver(const std::string &critic_solver);
  void LoadActorWeights(const std::string &actor_model_file);
  void LoadCriticWeights(const std::string &critic_weights);
  void LoadReplayMemory(const std::string &filename);

  // Snapshot the model/solver/replay memory (dqn.cpp:582-620)
  void Snapshot();
  void Snapshot(const std::string &snapshot_prefix, bool remove_old = false, bool snapshot_memory = true);

  ActorOutput GetRandomActorOutput();                                   // dqn.cpp:664-682
  ActorOutput SelectAction(const InputStates &input_states, double epsilon);   // dqn.cpp:684-686
  std::vector<ActorOutput> SelectActions(const std::vector<InputStates> &states_batch, double epsilon);  // :695-711
  Action SampleAction(const ActorOutput &actor_output);                 // dqn.cpp:180-194
  float EvaluateAction(const InputStates &input_states, const ActorOutput &action);  // dqn.cpp:688-693

  void AddTransition(const Transition &transition);                     // dqn.cpp:768-773
  void AddTransitions(const std::vector<Transition> &transitions);      // dqn.cpp:775-781
  void LabelTransitions(std::vector<Transition> &transitions);          // dqn.cpp:783-797
  void Update();                                                        // dqn.cpp:799-826

  void ClearReplayMemory();
  void SnapshotReplayMemory(const std::string &filename);               // dqn.cpp:1146-1178
  int memory_size() const;

  // Multi-agent sharing (dqn.cpp:1037-1083)
  void ShareLayer(caffe::Layer<float> &param_owner, caffe::Layer<float> &param_slave);
  void ShareParameters(DQN &other, int num_actor_layers_to_share, int num_critic_layers_to_share);
  void ShareReplayMemory(DQN &other);

  int min_iter() const { return std::min(actor_iter(), critic_iter()); }
  int max_iter() const { return std::max(actor_iter(), critic_iter()); }
  int critic_iter() const;
  int actor_iter() const;
  int state_size() const { return state_size_; }
  const std::string &save_path() const { return save_path_; }
  int unum() const { return unum_; }
  void set_unum(int unum) { unum_ = unum; }

  // Not in the reference: what the last Update() returned (dqn.cpp:971) and the raw handle.
  std::pair<float, float> last_update() const { return last_update_; }
  dqnb_handle_s *handle() { return h_; }

 protected:
  std::pair<float, float> UpdateActorCritic();                          // dqn.cpp:828-972
  std::vector<int> SampleTransitionsFromMemory(int n);                  // dqn.cpp:501-509
  void refresh_iters() const;

  caffe::SolverParameter actor_solver_param_, critic_solver_param_;
  const int replay_memory_capacity_;
  const double gamma_;
  dqnb_handle_s *h_;
  std::mt19937 random_engine;
  float smoothed_critic_loss_, smoothed_actor_loss_;
  int last_snapshot_iter_;
  std::string save_path_;
  const int state_size_;
  int batch_size_;
  int tid_;
  int unum_;
  mutable int actor_iter_cache_, critic_iter_cache_;
  mutable bool iters_dirty_;
  std::pair<float, float> last_update_;
  std::vector<int> hidden_;      // tower widths of both nets
  void snapshot_net(int net, const std::string &base) const;
  void load_weights(int net, const std::string &file);
  void restore_solver(int net, const std::string &file);
  void drain_pending();
  // Multi-agent sharing: every member of a group sees the first n layers of the group's nets as one set of weights
  // (ShareData aliases memory upstream; here the member that just updated writes its shared layers through to the others)
  struct ShareGroup;
  std::shared_ptr<ShareGroup> share_;
  // ShareReplayMemory (dqn.cpp:1081-1083) makes two DQNs point at ONE deque (replay_memory_ is a shared_ptr, dqn.hpp:187).
  // Every handle owns its ring in HBM, so the members of a replay group keep identical rings: what one member appends,
  // clears or loads is applied to every member's ring (same capacity, same eviction rule -> same contents), and each
  // samples its own copy.
  struct ReplayGroup;
  std::shared_ptr<ReplayGroup> replay_share_;
  std::vector<DQN *> replay_targets();          // this object, or every member of its replay group
  void add_rows(int n, const float *s, const float *a, const float *r, const float *mc, const float *sn, const uint8_t *term);
  // A handle is thread-compatible, not thread-safe (like a Caffe net).  Upstream every DQN is used by its own agent
  // thread only - except that thread 0 reaches into its teammates' objects to share layers / replay memory while they
  // are already playing (dqn_main.cpp:305-323) and, here, a member's update writes its shared layers into the others'
  // handles.  Every method that touches the handle therefore holds the object's mutex (recursive: Update -> Snapshot).
  mutable std::recursive_mutex mu_;
  long long pending_step_ = 0;   // -async_update: sequence number of the update whose results are still to be read
};

caffe::NetParameter CreateActorNet(int state_size);    // dqn.cpp:418-429
caffe::NetParameter CreateCriticNet(int state_size);   // dqn.cpp:431-454
// <prefix>_{actor,critic}.prototxt (dqn_main.cpp:232-246), protobuf text format: what WriteProtoToTextFile
// emits for the nets above / a parser for that family of files (tower widths and state size are honoured,
// anything the kernels do not implement aborts).  prototxt.cpp
std::string NetPrototxt(const caffe::NetParameter &np, int batch_size = kMinibatchSize);
void WriteNetPrototxt(const caffe::NetParameter &np, const std::string &filename, int batch_size = kMinibatchSize);
void ParseNetPrototxtOrDie(const std::string &text, const std::string &origin, bool critic, caffe::NetParameter *np);
void ReadNetPrototxtOrDie(const std::string &filename, bool critic, caffe::NetParameter *np);

Action GetAction(const ActorOutput &actor_output);     // dqn.cpp:196-208
std::vector<std::string> FilesMatchingRegexp(const std::string &regexp);   // dqn.cpp:559-580
void RemoveFilesMatchingRegexp(const std::string &regexp);
void RemoveSnapshots(const std::string &regexp, int min_iter);
void FindLatestSnapshot(const std::string &snapshot_prefix, std::string &actor_snapshot,
                        std::string &critic_snapshot, std::string &memory_snapshot);
int FindHiScore(const std::string &snapshot_prefix);
std::string PrintActorOutput(const ActorOutput &actor_output);

}  // namespace dqn

#endif /* DQN_HPP_ */
