// hfo_game.cpp — environment glue and reward shaping, restating the behaviour of the reference's
// src/hfo_game.cpp (scalar per-step host code; not on the GPU hot path).
#include "hfo_game.hpp"

#include <algorithm>
#include <cmath>

#include "shim/flags.hpp"
#include "shim/logging.hpp"

using namespace hfo;

DEFINE_string(server_cmd, "./bin/HFO --fullstate --frames-per-trial 500", "Command executed to start the HFO server.");
DEFINE_string(config_dir, "bin/formations-dt", "Directory containing HFO config files.");
DEFINE_string(server_addr, "localhost", "Address of rcssserver.");
DEFINE_string(team_name, "base_left", "Name of team for agents.");
DEFINE_bool(play_goalie, false, "Should the agent play goalie.");
DEFINE_string(record_dir, "", "Directory to record states,actions,rewards.");

// hfo_game.cpp:22-67: these spawn rcssserver / NPC processes upstream; nothing to spawn here.
void StartHFOServer(int port, int, int, int, int) { LOG(INFO) << "in-process HFO stand-in: no server to start (port " << port << ")"; }
void StartDummyTeammate(int) {}
void StartDummyGoalie(int) {}
void StartChaser(int, std::string, bool) {}
void StopHFOServer() {}

void ConnectToServer(hfo::HFOEnvironment &hfo_env, int port) {  // hfo_game.cpp:69-78
  hfo_env.connectToServer(LOW_LEVEL_FEATURE_SET, FLAGS_config_dir, port, FLAGS_server_addr, FLAGS_team_name,
                          FLAGS_play_goalie, FLAGS_record_dir);
}

Action GetRandomHFOAction(std::mt19937 &random_engine) {  // hfo_game.cpp:80-107: same draw order
  const action_t a = (action_t)std::uniform_int_distribution<int>(DASH, KICK)(random_engine);
  auto U = [&](float lo, float hi) { return std::uniform_real_distribution<float>(lo, hi)(random_engine); };
  Action act = {a, 0.f, 0.f};
  if (a == DASH) { act.arg1 = U(-100.0, 100.0); act.arg2 = U(-180.0, 180.0); }
  else if (a == TURN || a == TACKLE) { act.arg1 = U(-180.0, 180.0); }
  else if (a == KICK) { act.arg1 = U(0.0, 100.0); act.arg2 = U(-180.0, 180.0); }
  else LOG(FATAL) << "Invalid Action Index: " << a;
  return act;
}

HFOGameState::HFOGameState(int unum)
    : old_ball_prox(0), ball_prox_delta(0), old_kickable(0), kickable_delta(0), old_ball_dist_goal(0),
      ball_dist_goal_delta(0), steps(0), total_reward(0), extrinsic_reward(0), status(IN_GAME),
      episode_over(false), got_kickable_reward(false), our_unum(unum), pass_active(false) {}
HFOGameState::~HFOGameState() {}

static float signed_angle(float sin_v, float cos_v) {
  float a = std::acos(std::max(-1.f, std::min(1.f, cos_v)));
  return sin_v < 0 ? -a : a;
}

void HFOGameState::update(HFOEnvironment &hfo) {  // hfo_game.cpp:122-173
  status = hfo.step();
  if (status == SERVER_DOWN) LOG(FATAL) << "Server Down!";
  if (status != IN_GAME) episode_over = true;
  const std::vector<float> &f = hfo.getState();
  const float ball_proximity = f[53], goal_proximity = f[15], kickable = f[12];
  const float ball_dist = 1.0f - ball_proximity, goal_dist = 1.0f - goal_proximity;
  const float ball_ang = signed_angle(f[51], f[52]), goal_ang = signed_angle(f[13], f[14]);
  const float alpha = std::max(ball_ang, goal_ang) - std::min(ball_ang, goal_ang);
  // law of cosines: distance between ball and goal
  const float ball_dist_goal = std::sqrt(ball_dist * ball_dist + goal_dist * goal_dist - 2. * ball_dist * goal_dist * std::cos(alpha));
  if (f[54] && f[55] > kPassVelThreshold) pass_active = true;
  if (steps > 0) {
    ball_prox_delta = ball_proximity - old_ball_prox;
    kickable_delta = kickable - old_kickable;
    ball_dist_goal_delta = ball_dist_goal - old_ball_dist_goal;
  }
  old_ball_prox = ball_proximity; old_kickable = kickable; old_ball_dist_goal = ball_dist_goal;
  if (episode_over) { ball_prox_delta = 0; kickable_delta = 0; ball_dist_goal_delta = 0; }
  old_player_on_ball = player_on_ball;
  player_on_ball = hfo.playerOnBall();
  steps++;
}

float HFOGameState::reward() {  // hfo_game.cpp:175-186 (pass_reward is computed but not added upstream)
  const float mtb = move_to_ball_reward();
  const float ktg = 3. * kick_to_goal_reward();
  const float pass = 3. * pass_reward();
  (void)pass;
  const float eot = EOT_reward();
  const float r = mtb + ktg + eot;
  extrinsic_reward += eot;
  total_reward += r;
  return r;
}

float HFOGameState::move_to_ball_reward() {  // hfo_game.cpp:190-200
  float r = 0;
  if (player_on_ball.unum < 0 || player_on_ball.unum == our_unum) r += ball_prox_delta;
  if (kickable_delta >= 1 && !got_kickable_reward) { r += 1.0; got_kickable_reward = true; }
  return r;
}

float HFOGameState::kick_to_goal_reward() {  // hfo_game.cpp:203-210
  if (player_on_ball.unum == our_unum) return -ball_dist_goal_delta;
  if (got_kickable_reward) return 0.2 * -ball_dist_goal_delta;
  return 0;
}

float HFOGameState::EOT_reward() {  // hfo_game.cpp:212-227
  if (status == GOAL) {
    CHECK(old_player_on_ball.side == LEFT) << "Unexpected side: " << old_player_on_ball.side;
    return player_on_ball.unum == our_unum ? 5 : 1;
  }
  return 0;
}

float HFOGameState::pass_reward() {  // hfo_game.cpp:229-236
  if (pass_active && player_on_ball.unum > 0 && player_on_ball.unum != old_player_on_ball.unum) {
    pass_active = false;
    return 1;
  }
  return 0;
}
