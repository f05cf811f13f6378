// hfo_game.hpp — mirror of the reference's src/hfo_game.hpp (API kept name-for-name, lines 7-60):
// Action, NumStateFeatures, server/NPC launchers, ConnectToServer, GetRandomHFOAction, HFOGameState.
// The launchers are no-ops here: there is no rcssserver in the image, the environment behind
// <HFO.hpp> is the in-process stand-in (shim/HFO.hpp).
#ifndef HFO_GAME_HPP_
#define HFO_GAME_HPP_

#include <HFO.hpp>
#include <random>
#include <string>

struct Action {
  hfo::action_t action;
  float arg1;
  float arg2;
};

inline int NumStateFeatures(int num_players) { return 50 + 9 * num_players; }  // hfo_game.hpp:14-16

constexpr auto kPassVelThreshold = -.5;

void StartHFOServer(int port, int offense_agents, int offense_npcs, int defense_agents, int defense_npcs);
void StartDummyTeammate(int port);
void StartDummyGoalie(int port);
void StartChaser(int port, std::string team_name, bool goalie);
void StopHFOServer();
void ConnectToServer(hfo::HFOEnvironment &hfo_env, int port = 6000);
Action GetRandomHFOAction(std::mt19937 &random_engine);

class HFOGameState {
 public:
  explicit HFOGameState(int unum);
  ~HFOGameState();
  void update(hfo::HFOEnvironment &hfo);
  float reward();
  float move_to_ball_reward();
  float kick_to_goal_reward();
  float EOT_reward();
  float pass_reward();

 public:
  float old_ball_prox, ball_prox_delta;
  float old_kickable, kickable_delta;
  float old_ball_dist_goal, ball_dist_goal_delta;
  int steps;
  double total_reward;
  double extrinsic_reward;
  hfo::status_t status;
  bool episode_over;
  bool got_kickable_reward;
  hfo::Player old_player_on_ball, player_on_ball;
  int our_unum;
  bool pass_active;
};

#endif /* HFO_GAME_HPP_ */
