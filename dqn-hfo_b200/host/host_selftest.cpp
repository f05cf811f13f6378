// host_selftest.cpp — CPU-only checks of the host mirror (no GPU, no libdqn_b200 calls): flag
// parsing, GetAction semantics, LabelTransitions arithmetic, the in-process HFO stand-in and the
// reward shaping of HFOGameState.  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

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
  {  // <prefix>_{actor,critic}.prototxt (dqn_main.cpp:232-246): what we write is what we read ...
    shim::set_flag("hidden", "1024,512,256,128");
    for (int critic = 0; critic < 2; ++critic) {
      const caffe::NetParameter np = critic ? dqn::CreateCriticNet(59) : dqn::CreateActorNet(59);
      const std::string text = dqn::NetPrototxt(np, 32);
      EXPECT(text.find("name: \"" + std::string(critic ? "Critic" : "Actor") + "\"") == 0);
      EXPECT(text.find("force_backward: true") != std::string::npos);
      EXPECT(text.find("height: 59") != std::string::npos);
      EXPECT(text.find("name: \"ip4_relu_layer\"") != std::string::npos && text.find("negative_slope: 0.01") != std::string::npos);
      EXPECT((text.find("name: \"q_values_layer\"") != std::string::npos) == (critic == 1));
      EXPECT((text.find("name: \"actionpara_layer\"") != std::string::npos) == (critic == 0));
      EXPECT((text.find("type: \"Concat\"") != std::string::npos) == (critic == 1));
      caffe::NetParameter back;
      dqn::ParseNetPrototxtOrDie(text, "roundtrip", critic != 0, &back);
      EXPECT(back.hidden == np.hidden && back.state_size == 59 && back.critic == (critic != 0) && back.name() == np.name());
    }
    // ... and a hand-edited file of the same family (comments, other widths and depth, single-line messages,
    // the `field: { }` spelling) gives the tower the user asked for
    const char *edited =
        "# a narrower actor\n"
        "name: \"Actor\"  force_backward: true\n"
        "layer { name: \"state_input_layer\" type: \"MemoryData\" top: \"states\" top: \"dummy1\"\n"
        "        memory_data_param { batch_size: 32 channels: 1 height: 77 width: 1 } }\n"
        "layer { name: \"silence\" type: \"Silence\" bottom: \"dummy1\" }\n"
        "layer { name: \"ip1_layer\" type: \"InnerProduct\" bottom: \"states\" top: \"ip1\"\n"
        "        inner_product_param: { num_output: 300 weight_filler { type: \"gaussian\" std: 0.01 } } }\n"
        "layer { name: \"ip1_relu_layer\" type: \"ReLU\" bottom: \"ip1\" top: \"ip1\" relu_param { negative_slope: 0.01 } }\n"
        "layer { name: \"ip2_layer\" type: \"InnerProduct\" bottom: \"ip1\" top: \"ip2\" inner_product_param { num_output: 200 } }\n"
        "layer { name: \"ip2_relu_layer\" type: \"ReLU\" bottom: \"ip2\" top: \"ip2\" relu_param { negative_slope: 1e-2 } }\n"
        "layer { name: \"action_layer\" type: \"InnerProduct\" bottom: \"ip2\" top: \"actions\" inner_product_param { num_output: 4 } }\n"
        "layer { name: \"actionpara_layer\" type: \"InnerProduct\" bottom: \"ip2\" top: \"action_params\" inner_product_param { num_output: 6 } }\n";
    caffe::NetParameter ed;
    dqn::ParseNetPrototxtOrDie(edited, "edited", false, &ed);
    EXPECT(ed.hidden.size() == 2 && ed.hidden[0] == 300 && ed.hidden[1] == 200 && ed.state_size == 77 && !ed.critic);
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
