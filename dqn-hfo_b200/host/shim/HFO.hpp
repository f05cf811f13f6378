// HFO.hpp — stand-in for libhfo's header (HFO and rcssserver are not in this image).  Same enums
// and HFOEnvironment method names as the real interface the reference calls (dqn_main.cpp:97-153,
// hfo_game.cpp:69-78,:122-173); behind it sits a small deterministic 2-D half-field model so that
// rollouts, reward shaping and the training loop run in-process with no server.
#pragma once
#include <random>
#include <string>
#include <vector>

namespace hfo {

enum action_t { DASH, TURN, TACKLE, KICK, KICK_TO, MOVE_TO, DRIBBLE_TO, INTERCEPT, MOVE, SHOOT, PASS, DRIBBLE, CATCH, NOOP, QUIT };
enum status_t { IN_GAME, GOAL, CAPTURED_BY_DEFENSE, OUT_OF_BOUNDS, OUT_OF_TIME, SERVER_DOWN };
enum side_t { RIGHT = -1, NEUTRAL = 0, LEFT = 1 };
enum feature_set_t { LOW_LEVEL_FEATURE_SET, HIGH_LEVEL_FEATURE_SET };

struct Player { side_t side = NEUTRAL; int unum = -1; };

std::string ActionToString(action_t a);
std::string StatusToString(status_t s);

class HFOEnvironment {
 public:
  HFOEnvironment();
  void connectToServer(feature_set_t feature_set = LOW_LEVEL_FEATURE_SET, std::string config_dir = "",
                       int server_port = 6000, std::string server_addr = "localhost",
                       std::string team_name = "base_left", bool play_goalie = false,
                       std::string record_dir = "");
  const std::vector<float> &getState();
  void act(action_t action, float arg1 = 0.f, float arg2 = 0.f);
  status_t step();
  Player playerOnBall();
  int getUnum() { return unum_; }
  int getNumTeammates() { return 0; }
  int getNumOpponents() { return 0; }
  // fake-environment knobs (not part of libhfo)
  void configure(int num_features, int frames_per_trial, unsigned seed);

 private:
  void reset_episode();
  void fill_features();
  int num_features_ = 59, frames_per_trial_ = 500, unum_ = 11, frame_ = 0;
  bool fresh_ = true, configured_ = false;
  float px_, py_, heading_, bx_, by_, bvx_, bvy_;
  action_t pending_ = NOOP;
  float arg1_ = 0.f, arg2_ = 0.f;
  std::mt19937 rng_;
  std::vector<float> feat_;
  Player on_ball_;
};

}  // namespace hfo
