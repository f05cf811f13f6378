// host_selftest.cpp — CPU-only checks of the host mirror (no GPU, no libdqn_b200 calls): flag
// parsing, GetAction semantics, LabelTransitions arithmetic, the in-process HFO stand-in and the
// reward shaping of HFOGameState.  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "dqn.hpp"
#include "hfo_game.hpp"
#include "shim/flags.hpp"
#include "shim/logging.hpp"

namespace dqn {
DECLARE_double(gamma);
DECLARE_int32(batch_size);
DECLARE_string(hidden);
}
using dqn::FLAGS_batch_size;
using dqn::FLAGS_gamma;
using dqn::FLAGS_hidden;

static int failures = 0;
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

int main() {
  {  // flags
    const char *av[] = {"prog", "-gamma=0.5", "--batch_size", "1024", "-hidden=64,32", "positional"};
    int ac = 6; char **a = const_cast<char **>(av);
    shim::ParseCommandLineFlags(&ac, &a, true);
    EXPECT(FLAGS_gamma == 0.5); EXPECT(FLAGS_batch_size == 1024); EXPECT(FLAGS_hidden == "64,32");
    EXPECT(ac == 2 && std::strcmp(a[1], "positional") == 0);
    shim::set_flag("gamma", "0.99"); shim::set_flag("batch_size", "32"); shim::set_flag("hidden", "1024,512,256,128");
  }
  {  // GetAction (dqn.cpp:196-208): tackle masked, first max wins, param offsets
    dqn::ActorOutput o = {0.1f, 0.5f, 9.0f, 0.5f, 10, 20, 30, 40, 50, 60};
    Action a = dqn::GetAction(o);
    EXPECT(a.action == hfo::TURN && a.arg1 == 30.f && a.arg2 == 0.f);
    o[3] = 0.6f; a = dqn::GetAction(o);
    EXPECT(a.action == hfo::KICK && a.arg1 == 50.f && a.arg2 == 60.f);
    o[0] = 0.7f; a = dqn::GetAction(o);
    EXPECT(a.action == hfo::DASH && a.arg1 == 10.f && a.arg2 == 20.f);
    EXPECT(dqn::PrintActorOutput(o).find("Dash(10.000000, 20.000000)=0.700000") == 0);
  }
  {  // NumStateFeatures (hfo_game.hpp:14-16)
    EXPECT(NumStateFeatures(1) == 59 && NumStateFeatures(3) == 77);
  }
  {  // in-process environment: deterministic, bounded, ends episodes; reward shaping is finite
    for (int rep = 0; rep < 2; ++rep) {
      hfo::HFOEnvironment env;
      env.configure(59, 50, 7);
      static double first_total = 0;
      double total = 0;
      int episodes = 0, steps = 0;
      std::mt19937 rng(3);
      for (int e = 0; e < 5; ++e) {
        HFOGameState game(env.getUnum());
        env.act(hfo::DASH, 0, 0);
        game.update(env);
        EXPECT(!game.episode_over);
        while (!game.episode_over) {
          const std::vector<float> &f = env.getState();
          EXPECT((int)f.size() == 59);
          for (float v : f) EXPECT(std::isfinite(v) && v >= -1.0001f && v <= 1.0001f);
          Action a = GetRandomHFOAction(rng);
          env.act(a.action, a.arg1, a.arg2);
          game.update(env);
          const float r = game.reward();
          EXPECT(std::isfinite(r));
          ++steps;
        }
        EXPECT(game.steps <= 51);
        total += game.total_reward;
        ++episodes;
      }
      EXPECT(episodes == 5 && steps > 5);
      if (rep == 0) first_total = total; else EXPECT(total == first_total);
    }
  }
  {  // GetRandomHFOAction ranges (hfo_game.cpp:80-107)
    std::mt19937 rng(1);
    for (int i = 0; i < 1000; ++i) {
      Action a = GetRandomHFOAction(rng);
      EXPECT(a.action >= hfo::DASH && a.action <= hfo::KICK);
      if (a.action == hfo::KICK) EXPECT(a.arg1 >= 0.f && a.arg1 <= 100.f);
      EXPECT(std::fabs(a.arg2) <= 180.f);
    }
  }
  std::printf(failures ? "host_selftest: %d FAILURES\n" : "host_selftest: ok\n", failures);
  return failures ? 1 : 0;
}
