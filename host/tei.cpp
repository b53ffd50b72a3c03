// tei.cpp -- the reference's TEI engine front-end (tei/src/main.rs:63-300, tei/src/protocol.rs) as a C++
// host over libtakzero_b200.so.  One search tree (game 0 of the handle) searched with
// Node::simulate_batch(net, env, BETA = 0, BATCH_SIZE = 128) (main.rs:253), tree reuse through
// Node::descend when the new `position` extends the previous one (main.rs:175-184), `info` lines in the
// format of protocol.rs:240-273, `bestmove` = select_best_action.
// Differences: board size / komi / model are flags (the reference is compiled for one size and takes the
// model through a TEI option); stdin is polled between batches instead of being read by a second thread.
#include <poll.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../include/takzero_b200.hpp"

using namespace takzero;
using Clock = std::chrono::steady_clock;

static const int BATCH_SIZE = 128;  // tei/src/main.rs:26
static const float BETA = 0.0f;     // tei/src/main.rs:27

static bool stdin_ready() {
    pollfd p{STDIN_FILENO, POLLIN, 0};
    return poll(&p, 1, 0) > 0;
}

// protocol.rs:240-273
static std::string info_line(long long ms, unsigned long long nodes, const tz_root_t& root, const std::vector<Move>& pv) {
    Eval score;
    score.tag = root.eval_tag;
    score.ply = root.eval_bits;
    const float s = score.to_f32();
    std::ostringstream o;
    o << "info time " << ms << " nodes " << nodes << " nps " << (ms > 0 ? 1000 * nodes / (unsigned long long)ms : 0);
    switch (score.tag) {
        case 1: o << " wdl 1000 0 0"; break;
        case 2: o << " wdl 0 0 1000"; break;
        case 3: o << " wdl 0 1000 0"; break;
        default: {
            const int per_mille = 500 + (int)std::lround(s * 500.0f);
            o << " wdl " << per_mille << " 0 " << 1000 - per_mille;
        }
    }
    if (score.tag == 1) o << " score mate " << (score.ply + 1) / 2;
    if (score.tag == 2) o << " score mate -" << (score.ply + 1) / 2;
    o << " score cp " << (int)std::lround(s * 100.0f) << " pv";
    for (Move m : pv) o << ' ' << move_to_string(m);
    return o.str();
}

int main(int argc, char** argv) {
    int board = 6, half_komi = 4, device = 0;
    std::string weights;
    unsigned arena_slots = 1u << 22;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        const char* v = argv[i + 1];
        if (k == "--board") board = std::atoi(v);
        else if (k == "--half-komi") half_komi = std::atoi(v);
        else if (k == "--device") device = std::atoi(v);
        else if (k == "--weights") weights = v;
        else if (k == "--arena-slots") arena_slots = (unsigned)std::atoi(v);
    }
    try {
        // one game (the tree) with a large arena; the evaluation queue holds BATCH_SIZE leaves
        BatchedMCTS mcts(board, half_komi, 1, device, 0, arena_slots, BATCH_SIZE);
        if (!weights.empty()) {
            mcts.load_model(weights);
            mcts.set_agent(TZ_AGENT_NETWORK);
        }
        tz_state_t start_env = mcts.envs()[0];  // Env::default()
        mcts.tree_simulate_simple(0.0f);         // main.rs:139-141
        tz_state_t env = start_env;
        std::string last_position = "startpos";
        std::vector<Move> last_moves;
        auto restart = [&](const tz_state_t& e) {
            env = e;
            mcts.set_positions(std::vector<tz_state_t>(1, e));  // node = Node::default()
        };
        std::string line;
        std::deque<std::string> pending;  // commands that arrived while a search was running
        auto next_line = [&](std::string& out) {
            if (!pending.empty()) {
                out = pending.front();
                pending.pop_front();
                return true;
            }
            return (bool)std::getline(std::cin, out);
        };
        while (next_line(line)) {
            std::istringstream in(line);
            std::string cmd;
            in >> cmd;
            if (cmd == "tei") {
                std::cout << "id name takzero_b200\nid author takzero_b200 builders\n";
                std::cout << "option name HalfKomi type spin default " << half_komi << " min " << half_komi << " max "
                          << half_komi << "\nteiok" << std::endl;
            } else if (cmd == "isready") {
                std::cout << "readyok" << std::endl;
            } else if (cmd == "teinewgame") {
                int size = 0;
                in >> size;
                if (size != board) {
                    std::cerr << "the engine is started only for size " << board << std::endl;
                    return 1;
                }
                restart(start_env);
                last_position = "startpos";
                last_moves.clear();
            } else if (cmd == "position") {
                std::string kind, tok, position;
                in >> kind;
                tz_state_t base = start_env;
                if (kind == "startpos") {
                    position = "startpos";
                    in >> tok;  // "moves" or nothing
                } else {  // tps <board> <player> <move number> [moves ...]
                    std::string a, b, c;
                    in >> a >> b >> c;
                    position = a + " " + b + " " + c;
                    if (!parse_tps(position, board, &base)) {
                        std::cerr << "bad tps" << std::endl;
                        continue;
                    }
                    in >> tok;
                }
                std::vector<Move> moves;
                bool ok = true;
                while (in >> tok) {
                    Move m;
                    if (!parse_move(tok, &m)) {
                        std::cerr << "could not parse move " << tok << std::endl;
                        ok = false;
                        break;
                    }
                    moves.push_back(m);
                }
                if (!ok) continue;
                const bool extends = position == last_position && moves.size() >= last_moves.size() &&
                                     std::equal(last_moves.begin(), last_moves.end(), moves.begin());
                if (extends) {  // tree re-use (main.rs:175-184)
                    for (size_t i = last_moves.size(); i < moves.size(); i++) mcts.tree_descend(moves[i]);
                    env = mcts.envs()[0];
                } else {
                    tz_state_t e = base;
                    for (Move m : moves) {
                        int played = 0;
                        check(tz_apply(mcts.handle(), &e, &m, 1, &played));
                        if (!played) {
                            std::cerr << "could not play move " << move_to_string(m) << std::endl;
                            break;
                        }
                    }
                    restart(e);
                }
                last_position = position;
                last_moves = moves;
            } else if (cmd == "go") {
                unsigned long long nodes = 0;
                long long move_time_ms = -1;
                std::string opt;
                while (in >> opt) {
                    if (opt == "nodes") in >> nodes;
                    else if (opt == "movetime") in >> move_time_ms;
                    else if (opt == "infinite") nodes = ~0ull;
                }
                if (nodes == 0 && move_time_ms < 0) std::cerr << "no understood stopping condition given" << std::endl;
                const unsigned visits_at_start = mcts.root_stats()[0].visit_count;
                const auto start = Clock::now();
                auto last_info = start;
                bool sent_info = false;
                auto elapsed_ms = [&]() {
                    return (long long)std::chrono::duration_cast<std::chrono::milliseconds>(Clock::now() - start).count();
                };
                tz_root_t root{};
                while (true) {
                    mcts.tree_simulate_batch(BETA, BATCH_SIZE);
                    root = mcts.root_stats()[0];
                    const unsigned long long visits = root.visit_count - visits_at_start;
                    const bool done = (nodes && visits >= nodes) || (move_time_ms >= 0 && elapsed_ms() >= move_time_ms) ||
                                      (nodes == 0 && move_time_ms < 0);
                    if (std::chrono::duration_cast<std::chrono::milliseconds>(Clock::now() - last_info).count() >= 300) {
                        std::cout << info_line(elapsed_ms(), visits, root, mcts.principal_variation()) << std::endl;
                        sent_info = true;
                        last_info = Clock::now();
                    }
                    if (done) break;
                    if (stdin_ready()) {  // `stop` / `quit` end the search (GoStatus::Stopping); the rest waits
                        std::string peek;
                        if (!std::getline(std::cin, peek)) break;
                        if (peek.rfind("stop", 0) == 0) break;
                        if (peek.rfind("quit", 0) == 0) {
                            move_time_ms = -2;
                            break;
                        }
                        pending.push_back(peek);
                    }
                }
                const std::vector<Move> pv = mcts.principal_variation();
                if (!sent_info) std::cout << info_line(elapsed_ms(), root.visit_count - visits_at_start, root, pv) << std::endl;
                if (!pv.empty()) std::cout << "bestmove " << move_to_string(pv[0]) << std::endl;
                if (move_time_ms == -2) break;
            } else if (cmd == "stop") {
                // nothing is running
            } else if (cmd == "quit") {
                break;
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "tei: %s\n", e.what());
        return 1;
    }
    return 0;
}
