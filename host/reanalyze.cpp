// reanalyze.cpp -- the reference's `reanalyze` binary (reanalyze/src/main.rs:60-244) as a C++ host over
// libtakzero_b200.so: read `replays.txt`, expand every replay into its positions, sample a batch of fresh
// roots, search them with Gumbel sequential halving (beta = 0) and append `targets-reanalyze.txt` with
//   value  = root evaluation if solved, else -evaluation of the selected child   (main.rs:184-195)
//   policy = improved_policy(most_visited_count())                               (main.rs:196-202)
//   ube    = ube_target(0.25)                                                     (main.rs:203)
// Differences at the process boundary: constants are flags; positions are sampled WITH replacement by a
// counter-based hash of (seed, batch, slot) instead of `rand`'s `sample` (whose stream is not pinned); the
// model is a TZW1 file loaded once; no buffer-length throttle; stops after --batches iterations.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "../include/takzero_b200.hpp"

using namespace takzero;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

int main(int argc, char** argv) {
    std::string directory = ".", weights;
    int board = 6, half_komi = 4, games = 128, device = 0, batches = 1, sampled_actions = 64;
    unsigned budget = 768, arena_slots = 0;
    unsigned long long seed = 1;
    float ube_beta = 0.25f;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        if (i + 1 >= argc) {
            std::fprintf(stderr, "missing value for %s\n", k.c_str());
            return 2;
        }
        const char* v = argv[++i];
        if (k == "--directory") directory = v;
        else if (k == "--weights") weights = v;
        else if (k == "--board") board = std::atoi(v);
        else if (k == "--half-komi") half_komi = std::atoi(v);
        else if (k == "--games") games = std::atoi(v);
        else if (k == "--device") device = std::atoi(v);
        else if (k == "--batches") batches = std::atoi(v);
        else if (k == "--sampled-actions") sampled_actions = std::atoi(v);
        else if (k == "--budget") budget = (unsigned)std::atoi(v);
        else if (k == "--arena-slots") arena_slots = (unsigned)std::atoi(v);
        else if (k == "--seed") seed = std::strtoull(v, nullptr, 10);
        else {
            std::fprintf(stderr, "unknown flag %s\n", k.c_str());
            return 2;
        }
    }
    try {
        BatchedMCTS mcts(board, half_komi, games, device, 0, arena_slots);
        if (!weights.empty()) {
            mcts.set_weights(Weights::load(weights));
            mcts.set_agent(TZ_AGENT_NETWORK);
        }
        // fill_buffer_with_positions_from_replays (main.rs:262-285)
        std::vector<tz_state_t> positions;
        {
            std::ifstream f(directory + "/replays.txt");
            std::string line;
            while (std::getline(f, line)) {
                Replay r;
                if (!Replay::parse(line, board, &r)) continue;
                const std::vector<tz_state_t> st = mcts.replay_states(r);
                positions.insert(positions.end(), st.begin(), st.end());
            }
        }
        if ((int)positions.size() < games) {
            std::fprintf(stderr, "reanalyze: not enough positions yet (%zu)\n", positions.size());
            return 1;
        }
        const int stride = mcts.move_stride();
        const std::vector<float> zero_beta(games, 0.0f);
        for (int b = 0; b < batches; b++) {
            std::vector<tz_state_t> batch(games);
            for (int g = 0; g < games; g++)
                batch[g] = positions[mix64(seed * 0x9e3779b97f4a7c15ULL + (uint64_t)b * 1000003ULL + (uint64_t)g) % positions.size()];
            mcts.set_positions(batch);  // *node = Node::default(); *env = replay_env
            const std::vector<Move> selected = mcts.gumbel_sequential_halving(zero_beta, sampled_actions, budget, seed + b);
            const std::vector<tz_root_t> roots = mcts.root_stats();
            const BatchedMCTS::Children ch = mcts.root_children();
            const BatchedMCTS::RootTargets rt = mcts.targets(-1.0f, ube_beta);
            std::string contents;
            for (int g = 0; g < games; g++) {
                Eval value;
                if (roots[g].eval_tag != 0) {
                    value.tag = roots[g].eval_tag;
                    value.ply = roots[g].eval_bits;
                } else {
                    int found = -1;
                    for (int i = 0; i < ch.n[g]; i++)
                        if (ch.moves[(size_t)g * stride + i] == selected[g]) {
                            found = i;
                            break;
                        }
                    if (found < 0) throw std::runtime_error("all non-terminal nodes should have at least one child");
                    Eval child;
                    child.tag = ch.eval_tag[(size_t)g * stride + found];
                    child.ply = ch.eval_bits[(size_t)g * stride + found];
                    value = child.negate();
                }
                Target t;
                t.env = batch[g];
                for (int i = 0; i < rt.n[g]; i++)
                    t.policy.emplace_back(rt.moves[(size_t)g * stride + i], rt.policy[(size_t)g * stride + i]);
                t.value = value.to_f32();
                t.ube = rt.ube[g];
                contents += t.to_string(board);
            }
            std::ofstream out(directory + "/targets-reanalyze.txt", std::ios::app | std::ios::binary);
            if (!out || !(out << contents))
                std::fprintf(stderr, "Could not save targets to file, so here they are instead:\n%s", contents.c_str());
        }
        std::printf("reanalyze: %d batches of %d positions out of %zu\n", batches, games, positions.size());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "reanalyze: %s\n", e.what());
        return 1;
    }
    return 0;
}
