// dqn_main.cpp — the caller side of the drop-in: same flags, epsilon schedule, episode loop,
// update_ratio bookkeeping, evaluation and log lines as the reference's src/dqn_main.cpp, written
// against the dqn.hpp / hfo_game.hpp mirrors.  One learning agent per process here; the reference's
// multi-thread agent sharing (dqn_main.cpp:293-323) is out of the hot-path scope.
#include <HFO.hpp>
#include <algorithm>
#include <cmath>
#include <deque>
#include <functional>
#include <iostream>
#include <limits>
#include <tuple>

#include <sys/stat.h>

#include "dqn.hpp"
#include "hfo_game.hpp"
#include "shim/flags.hpp"
#include "shim/logging.hpp"

using namespace hfo;

DEFINE_bool(gpu, true, "Kept for command-line compatibility: this build always runs on the GPU");
DEFINE_bool(benchmark, false, "Benchmark the network and exit");
DEFINE_bool(learn_offline, false, "Just do updates on a fixed replaymemory.");
DEFINE_string(save, "", "Prefix for saving snapshots");
DEFINE_string(resume, "", "Prefix for resuming from. Default=save_path");
DEFINE_string(actor_weights, "", "The actor pretrained weights load (*.caffemodel).");
DEFINE_string(critic_weights, "", "The critic pretrained weights load (*.caffemodel).");
DEFINE_string(actor_snapshot, "", "The actor solver state to load (*.solverstate).");
DEFINE_string(critic_snapshot, "", "The critic solver state to load (*.solverstate).");
DEFINE_string(memory_snapshot, "", "The replay memory to load (*.replaymemory).");
DEFINE_string(solver, "Adam", "Solver Type.");
DEFINE_double(momentum, .95, "Solver momentum.");
DEFINE_double(momentum2, .999, "Solver momentum2.");
DEFINE_double(actor_lr, .00001, "Solver learning rate.");
DEFINE_double(critic_lr, .001, "Solver learning rate.");
DEFINE_double(clip_grad, 10, "Clip gradients.");
DEFINE_string(lr_policy, "fixed", "LR Policy.");
DEFINE_int32(max_iter, 10000000, "Custom max iter.");
DEFINE_int32(explore, 10000, "Iterations for epsilon to reach given value.");
DEFINE_double(epsilon, .1, "Value of epsilon after explore iterations.");
DEFINE_double(evaluate_with_epsilon, 0, "Epsilon value to be used in evaluation mode");
DEFINE_bool(evaluate, false, "Evaluation mode: only playing a game, no updates");
DEFINE_int32(evaluate_freq, 10000, "Frequency (steps) between evaluations");
DEFINE_int32(repeat_games, 100, "Number of games played in evaluation mode");
DEFINE_double(update_ratio, 0.1, "Ratio of new experiences to updates.");
DEFINE_int32(offense_agents, 1, "Number of agents playing offense");
DEFINE_int32(offense_npcs, 0, "Number of npcs playing offense");
DEFINE_int32(defense_agents, 0, "Number of agents playing defense");
DEFINE_int32(defense_npcs, 0, "Number of npcs playing defense");
DEFINE_int32(frames_per_trial, 500, "Episode length cap of the in-process environment (--frames-per-trial upstream)");
DEFINE_int32(benchmark_iters, 1000, "Updates timed by -benchmark (dqn.cpp:487 default)");
namespace dqn { DECLARE_int32(seed); DECLARE_int32(batch_size); }
using dqn::FLAGS_seed;

double CalculateEpsilon(const int iter) {  // dqn_main.cpp:65-71
  if (iter < FLAGS_explore) return 1.0 - (1.0 - FLAGS_epsilon) * (static_cast<double>(iter) / FLAGS_explore);
  return FLAGS_epsilon;
}

// Play one episode; returns (total reward, steps, final status, extrinsic reward)  (dqn_main.cpp:97-153)
std::tuple<double, int, status_t, double> PlayOneEpisode(HFOEnvironment &hfo, dqn::DQN &dqn, const double epsilon,
                                                         const bool update, const int tid) {
  (void)tid;
  std::vector<dqn::Transition> episode;
  HFOGameState game(dqn.unum());
  hfo.act(DASH, 0, 0);
  game.update(hfo);
  CHECK(!game.episode_over) << "Episode should not be over at beginning!";
  while (!game.episode_over) {
    const std::vector<float> &current_state = hfo.getState();
    CHECK_EQ((int)current_state.size(), dqn.state_size());
    dqn::InputStates input_states;
    input_states[0] = std::make_shared<dqn::StateData>(current_state);
    const dqn::ActorOutput actor_output = dqn.SelectAction(input_states, epsilon);
    VLOG(1) << "Step " << game.steps << " Actor_output: " << dqn::PrintActorOutput(actor_output);
    const Action action = dqn::GetAction(actor_output);
    VLOG(1) << "q_value: " << dqn.EvaluateAction(input_states, actor_output) << " Action: " << hfo::ActionToString(action.action);
    hfo.act(action.action, action.arg1, action.arg2);
    game.update(hfo);
    const float reward = game.reward();
    if (update) {
      const std::vector<float> &next_state = hfo.getState();
      CHECK_EQ((int)next_state.size(), dqn.state_size());
      if (game.status == IN_GAME)
        episode.push_back(dqn::Transition(input_states, actor_output, reward, 0, std::make_shared<dqn::StateData>(next_state)));
      else
        episode.push_back(dqn::Transition(input_states, actor_output, reward, 0, std::nullopt));
    }
  }
  if (update && !episode.empty()) {
    dqn.LabelTransitions(episode);
    dqn.AddTransitions(episode);
  }
  return std::make_tuple(game.total_reward, game.steps, game.status, game.extrinsic_reward);
}

template <class T>
static std::pair<double, double> get_avg_std(const std::vector<T> &data) {
  double sum = 0;
  for (const T &v : data) sum += v;
  const double avg = sum / static_cast<double>(data.size());
  double sq = 0;
  for (const T &v : data) sq += (v - avg) * (v - avg);
  return std::make_pair(avg, std::sqrt(sq / static_cast<double>(data.size() - 1)));
}

double Evaluate(HFOEnvironment &hfo, dqn::DQN &dqn, int tid) {  // dqn_main.cpp:171-204, same log line
  LOG(INFO) << "[Agent" << tid << "] Evaluating for " << FLAGS_repeat_games << " episodes with epsilon = " << FLAGS_evaluate_with_epsilon;
  std::vector<double> scores;
  std::vector<int> steps, successful_trial_steps;
  int goals = 0;
  for (int i = 0; i < FLAGS_repeat_games; ++i) {
    auto result = PlayOneEpisode(hfo, dqn, FLAGS_evaluate_with_epsilon, false, tid);
    scores.push_back(std::get<0>(result));
    steps.push_back(std::get<1>(result));
    if (std::get<2>(result) == GOAL) { goals++; successful_trial_steps.push_back(std::get<1>(result)); }
  }
  const auto score_dist = get_avg_std(scores), steps_dist = get_avg_std(steps), succ = get_avg_std(successful_trial_steps);
  const float goal_percent = goals / float(FLAGS_repeat_games);
  LOG(INFO) << "[Agent" << tid << "] Evaluation: actor_iter = " << dqn.actor_iter() << ", avg_reward = " << score_dist.first
            << ", reward_std = " << score_dist.second << ", avg_steps = " << steps_dist.first << ", steps_std = " << steps_dist.second
            << ", success_steps = " << succ.first << ", success_std = " << succ.second << ", goal_perc = " << goal_percent;
  return goal_percent;
}

void KeepPlayingGames(int tid, std::string save_prefix, int port) {  // dqn_main.cpp:206-382
  LOG(INFO) << "Thread " << tid << ", port=" << port << ", save_prefix=" << save_prefix;
  std::string actor_snapshot = FLAGS_actor_snapshot, critic_snapshot = FLAGS_critic_snapshot, memory_snapshot = FLAGS_memory_snapshot;
  const std::string resume = FLAGS_resume.empty() ? save_prefix : FLAGS_resume;
  if (actor_snapshot.empty() && critic_snapshot.empty() && memory_snapshot.empty())
    dqn::FindLatestSnapshot(resume, actor_snapshot, critic_snapshot, memory_snapshot);
  const int num_players = FLAGS_offense_agents + FLAGS_offense_npcs + FLAGS_defense_agents + FLAGS_defense_npcs;
  const int num_features = NumStateFeatures(num_players);
  caffe::SolverParameter actor_solver_param, critic_solver_param;   // dqn_main.cpp:247-262
  actor_solver_param.set_snapshot_prefix(save_prefix + "_actor");
  critic_solver_param.set_snapshot_prefix(save_prefix + "_critic");
  actor_solver_param.set_max_iter(FLAGS_max_iter); critic_solver_param.set_max_iter(FLAGS_max_iter);
  actor_solver_param.set_type(FLAGS_solver); critic_solver_param.set_type(FLAGS_solver);
  actor_solver_param.set_base_lr(FLAGS_actor_lr); critic_solver_param.set_base_lr(FLAGS_critic_lr);
  actor_solver_param.set_lr_policy(FLAGS_lr_policy); critic_solver_param.set_lr_policy(FLAGS_lr_policy);
  actor_solver_param.set_momentum(FLAGS_momentum); critic_solver_param.set_momentum(FLAGS_momentum);
  actor_solver_param.set_momentum2(FLAGS_momentum2); critic_solver_param.set_momentum2(FLAGS_momentum2);
  actor_solver_param.set_clip_gradients(FLAGS_clip_grad); critic_solver_param.set_clip_gradients(FLAGS_clip_grad);
  // dqn_main.cpp:232-246: the nets are described by <prefix>_{actor,critic}.prototxt; an existing file wins
  // (that is how a run's architecture is changed upstream), otherwise the default net is created and written
  struct stat st;
  const std::string actor_net_filename = save_prefix + "_actor.prototxt", critic_net_filename = save_prefix + "_critic.prototxt";
  const bool net_files = !FLAGS_save.empty();   // -benchmark / -evaluate may run without a save prefix: no files then
  if (net_files && stat(actor_net_filename.c_str(), &st) == 0 && S_ISREG(st.st_mode)) {
    dqn::ReadNetPrototxtOrDie(actor_net_filename, false, actor_solver_param.mutable_net_param());
  } else {
    *actor_solver_param.mutable_net_param() = dqn::CreateActorNet(num_features);
    if (net_files) dqn::WriteNetPrototxt(*actor_solver_param.mutable_net_param(), actor_net_filename, dqn::FLAGS_batch_size);
  }
  if (net_files && stat(critic_net_filename.c_str(), &st) == 0 && S_ISREG(st.st_mode)) {
    dqn::ReadNetPrototxtOrDie(critic_net_filename, true, critic_solver_param.mutable_net_param());
  } else {
    *critic_solver_param.mutable_net_param() = dqn::CreateCriticNet(num_features);
    if (net_files) dqn::WriteNetPrototxt(*critic_solver_param.mutable_net_param(), critic_net_filename, dqn::FLAGS_batch_size);
  }

  dqn::DQN *dqn = new dqn::DQN(actor_solver_param, critic_solver_param, save_prefix, num_features, tid);
  if (!actor_snapshot.empty()) dqn->RestoreActorSolver(actor_snapshot);
  else if (!FLAGS_actor_weights.empty()) dqn->LoadActorWeights(FLAGS_actor_weights);
  if (!critic_snapshot.empty()) dqn->RestoreCriticSolver(critic_snapshot);
  else if (!FLAGS_critic_weights.empty()) dqn->LoadCriticWeights(FLAGS_critic_weights);
  if (!memory_snapshot.empty()) dqn->LoadReplayMemory(memory_snapshot);

  HFOEnvironment env;
  env.configure(num_features, FLAGS_frames_per_trial, FLAGS_seed > 0 ? FLAGS_seed : 1);
  ConnectToServer(env, port);
  dqn->set_unum(env.getUnum());

  if (FLAGS_evaluate) { Evaluate(env, *dqn, tid); delete dqn; return; }
  if (FLAGS_benchmark) {   // dqn_main.cpp:332-338: one episode to seed the memory, then time updates
    PlayOneEpisode(env, *dqn, FLAGS_epsilon, true, tid);
    dqn->Benchmark(FLAGS_benchmark_iters);
    delete dqn;
    return;
  }
  if (FLAGS_learn_offline) {   // dqn_main.cpp:339-345
    while (dqn->max_iter() < FLAGS_max_iter) dqn->Update();
    dqn->Snapshot();
    delete dqn;
    return;
  }
  int last_eval_iter = dqn->max_iter(), episode = 0;
  double best_score = std::numeric_limits<double>::lowest();
  while (dqn->max_iter() < FLAGS_max_iter) {   // dqn_main.cpp:352-377
    const double epsilon = CalculateEpsilon(dqn->max_iter());
    auto result = PlayOneEpisode(env, *dqn, epsilon, true, tid);
    LOG(INFO) << "[Agent" << tid << "] Episode " << episode << " reward = " << std::get<0>(result);
    const int steps = std::get<1>(result);
    const int n_updates = int(steps * FLAGS_update_ratio);
    for (int i = 0; i < n_updates; ++i) dqn->Update();
    episode++;
    if (dqn->max_iter() >= last_eval_iter + FLAGS_evaluate_freq) {
      const double score = Evaluate(env, *dqn, tid);
      if (score > best_score) {
        LOG(INFO) << "[Agent" << tid << "] New High Score: " << score << ", actor_iter = " << dqn->actor_iter();
        best_score = score;
        dqn->Snapshot(save_prefix + "_HiScore" + std::to_string(int(score * 100)), false, false);
      }
      last_eval_iter = dqn->max_iter();
    }
  }
  dqn->Snapshot();
  delete dqn;
}

int main(int argc, char **argv) {  // dqn_main.cpp:388-442
  gflags::ParseCommandLineFlags(&argc, &argv, true);
  if (FLAGS_save.empty() && !FLAGS_evaluate && !FLAGS_benchmark) {
    LOG(ERROR) << "Save path (or evaluate) required but not set.";
    LOG(ERROR) << "Usage: " << argv[0] << " -save [path] or -evaluate / -benchmark";
    return 1;
  }
  const int port = (int)(std::hash<std::string>()(FLAGS_save) % 40000 + 20000);   // dqn_main.cpp:413-414
  StartHFOServer(port, FLAGS_offense_agents, FLAGS_offense_npcs, FLAGS_defense_agents, FLAGS_defense_npcs);
  KeepPlayingGames(0, FLAGS_save + "_agent0", port);
  return 0;
}
