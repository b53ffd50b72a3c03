// analysis.cpp -- the reference's `analysis` REPL (analysis/src/main.rs) as a C++ host over libtakzero_b200.so: one
// search tree (game 0 of a handle) that the user steers from stdin.  A line that parses as a move is played
// (`env.play` + `Node::descend`: the explored sub-tree is kept); any other line runs `simulate_batch(agent, env,
// BETA = 0, BATCH_SIZE = 128)`; after either, the root is printed with `impl Display for Node` (node/debug.rs).
// `--example` plays a whole game against itself: simulate, `select_best_action`, descend (main.rs:33-42).
// Differences at the process boundary: --model-path may be omitted (deterministic synthetic agent, for tests);
// board size and komi are flags; end of input ends the program.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "../include/takzero_b200.hpp"

using namespace takzero;

static const float BETA = 0.0f;

int main(int argc, char** argv) {
    int board = 6, half_komi = 4, device = 0, batch_size = 128;
    unsigned arena_slots = 1u << 22;
    std::string model_path, tps_text;
    bool example = false;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        if (k == "--example") {
            example = true;
            continue;
        }
        if (i + 1 >= argc) {
            std::fprintf(stderr, "missing value for %s\n", k.c_str());
            return 2;
        }
        const char* v = argv[++i];
        if (k == "--model-path") model_path = v;
        else if (k == "--tps") tps_text = v;
        else if (k == "--board") board = std::atoi(v);
        else if (k == "--half-komi") half_komi = std::atoi(v);
        else if (k == "--device") device = std::atoi(v);
        else if (k == "--batch-size") batch_size = std::atoi(v);
        else if (k == "--arena-slots") arena_slots = (unsigned)std::atoi(v);
        else {
            std::fprintf(stderr, "unknown flag %s\n", k.c_str());
            return 2;
        }
    }
    try {
        BatchedMCTS mcts(board, half_komi, 1, device, 0, arena_slots, batch_size);
        if (!model_path.empty()) {
            mcts.load_model(model_path);
            mcts.set_agent(TZ_AGENT_NETWORK);
        }
        tz_state_t env = mcts.envs()[0];  // Env::default()
        if (!tps_text.empty() && !parse_tps(tps_text, board, &env)) throw std::runtime_error("--tps is not valid TPS");
        mcts.set_positions(std::vector<tz_state_t>(1, env));  // node = Node::default()
        auto terminal = [&]() {
            int t = 0;
            check(tz_result(mcts.handle(), &env, 1, &t));
            return t;
        };
        if (example) {  // run_example (main.rs:33-42)
            while (terminal() == 0) {
                std::cout << "tps: " << tps(env, board) << "\n";
                mcts.tree_simulate_batch(BETA, batch_size);
                const Move action = mcts.select_best_actions()[0];
                std::cout << ">>> " << move_to_string(action) << std::endl;
                mcts.tree_descend(action);
                env = mcts.envs()[0];
            }
            return 0;
        }
        std::string input;
        for (;;) {
            std::cout << "tps: " << tps(env, board) << "\n>>> " << std::flush;
            if (!std::getline(std::cin, input)) break;
            const size_t a = input.find_first_not_of(" \t\r\n"), b = input.find_last_not_of(" \t\r\n");
            const std::string trim = a == std::string::npos ? "" : input.substr(a, b - a + 1);
            Move mov;
            if (parse_move(trim, &mov)) {
                tz_state_t next = env;
                int ok = 0;
                check(tz_apply(mcts.handle(), &next, &mov, 1, &ok));
                if (!ok) {
                    std::cerr << "illegal move " << trim << "\n";  // `env.play` error: the position stays
                    continue;
                }
                mcts.tree_descend(mov);
                env = mcts.envs()[0];
            } else {
                mcts.tree_simulate_batch(BETA, batch_size);
            }
            std::cout << node_display(mcts) << "\n";  // println!("{node}"): Display ends with a newline already
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "analysis: %s\n", e.what());
        return 1;
    }
    return 0;
}
