// evaluation.cpp -- the reference's `evaluation` binary (evaluation/src/main.rs:139-318) as a C++ host over
// libtakzero_b200.so: two networks play a batch of games against each other, each side searching its own trees.
//
// compete(white, black) (main.rs:222-318): two BatchedMCTS over the same starting positions, one per network.  The
// side to move searches with Gumbel sequential halving (beta 0), BOTH step with the selected moves (the other side's
// tree follows by `Node::descend`, i.e. keeps a sub-tree only if it had explored that move), finished games are
// collected from the searcher's `restart_terminal_envs` -- a game counts once, the first time its slot finishes --
// and the other side's slot is reset to the searcher's new position.  The result is from white's point of view;
// every pair of models plays each batch twice with colours swapped (main.rs:206-217) and prints
//   "{a} vs. {b}: Evaluation { wins: W, losses: L, draws: D } P%"      (parsed by python/get_match_results.py)
//
// Differences at the process boundary: constants are flags (BATCH_SIZE 64, MAX_MOVES 200, SAMPLED_ACTIONS 64,
// SEARCH_BUDGET 768); the match-up is given explicitly (--model-a / --model-b) or drawn from the sorted `*.ot` files
// of --model-path (every --step-th, `model_latest` skipped) by a counter-based hash instead of `rand`'s `sample`;
// openings come from --opening-book (TPS lines, a distinct sample) or `new_opening_with_random_steps` with 2..=3
// random plies; noise and restarts are drawn by the library from --seed; it stops after --rounds match-ups; without
// any model both sides use the deterministic synthetic agent (tests).
#include <dirent.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <unordered_set>
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

struct Evaluation {  // main.rs:322-350
    unsigned wins = 0, losses = 0, draws = 0;
    std::string to_string() const {  // `{:?}` of the struct followed by `{:.1}%` of the win rate
        char buf[160];
        const unsigned total = wins + losses + draws;
        if (total == 0) {
            std::snprintf(buf, sizeof(buf), "Evaluation { wins: %u, losses: %u, draws: %u } NaN%%", wins, losses, draws);
        } else {
            std::snprintf(buf, sizeof(buf), "Evaluation { wins: %u, losses: %u, draws: %u } %.1f%%", wins, losses, draws,
                          (double)wins / (double)total * 100.0);
        }
        return buf;
    }
};

struct Settings {
    int board = 4, half_komi = 4, games = 64, device = 0, max_moves = 200, sampled_actions = 64;
    unsigned budget = 768, arena_slots = 0;
    unsigned long long seed = 1;
    bool verbose = false;
};

static void use_model(BatchedMCTS& mcts, const std::string& path) {
    if (path.empty()) return;  // synthetic agent
    mcts.load_model(path);
    mcts.set_agent(TZ_AGENT_NETWORK);
}

// compete (main.rs:222-318); `round` only feeds the seeds
static Evaluation compete(BatchedMCTS& white, BatchedMCTS& black, const std::vector<tz_state_t>& games, const Settings& cfg,
                          uint64_t round) {
    Evaluation evaluation;
    const int G = cfg.games, n = cfg.board;
    white.set_positions(games);  // BatchedMCTS::from_envs
    black.set_positions(games);
    const std::vector<float> zero_beta(G, 0.0f);
    std::vector<uint8_t> done(G, 0);
    uint64_t ply = 0;
    for (int mv = 0; mv < cfg.max_moves; mv++) {
        for (int is_white = 1; is_white >= 0; is_white--, ply++) {
            if (std::all_of(done.begin(), done.end(), [](uint8_t d) { return d != 0; })) return evaluation;
            BatchedMCTS& current = is_white ? white : black;
            BatchedMCTS& other = is_white ? black : white;
            const uint64_t seed = cfg.seed + 1000003ULL * round + ply;
            const std::vector<Move> top = current.gumbel_sequential_halving(zero_beta, cfg.sampled_actions, cfg.budget, seed);
            current.step(top);
            other.step(top);
            const std::vector<tz_state_t> after = cfg.verbose ? current.envs() : std::vector<tz_state_t>();
            const std::vector<int> terminal = current.restart_terminal_envs(seed);
            std::vector<uint8_t> newly(G, 0);
            for (int g = 0; g < G; g++) {
                if (done[g] || terminal[g] == 0) continue;
                done[g] = newly[g] = 1;
                // the terminal is seen by the player to move AFTER the move: a loss for it is a win for the mover
                if (terminal[g] == 3) evaluation.draws++;
                else if ((terminal[g] == 2) == (is_white != 0)) evaluation.wins++;
                else evaluation.losses++;
                if (cfg.verbose) {
                    int result = 0;
                    check(tz_game_result(current.handle(), &after[g], 1, &result));
                    std::fprintf(stderr, "%s", current.finished_replay(g).to_string(n, result).c_str());
                }
            }
            // the other side's finished slots (every slot ever finished, like the reference's `.filter(done)`) take
            // the searcher's position with a fresh root
            if (std::any_of(done.begin(), done.end(), [](uint8_t d) { return d != 0; }))
                other.set_positions(current.envs(), done);
        }
    }
    return evaluation;
}

static std::vector<std::string> model_files(const std::string& dir, int step) {
    std::vector<std::string> out;
    if (DIR* d = opendir(dir.c_str())) {
        while (dirent* e = readdir(d)) {
            const std::string name = e->d_name;
            if (name.size() > 3 && name.compare(name.size() - 3, 3, ".ot") == 0 && name != "model_latest.ot") out.push_back(name);
        }
        closedir(d);
    }
    std::sort(out.begin(), out.end());
    std::vector<std::string> stepped;
    for (size_t i = 0; i < out.size(); i += (size_t)std::max(step, 1)) stepped.push_back(dir + "/" + out[i]);
    return stepped;
}

static std::string base_name(const std::string& path) {
    const size_t slash = path.find_last_of('/');
    return slash == std::string::npos ? path : path.substr(slash + 1);
}

int main(int argc, char** argv) {
    Settings cfg;
    std::string model_a, model_b, model_path, opening_book;
    int step = 1, rounds = 1;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        if (k == "--verbose") {
            cfg.verbose = true;
            continue;
        }
        if (i + 1 >= argc) {
            std::fprintf(stderr, "missing value for %s\n", k.c_str());
            return 2;
        }
        const char* v = argv[++i];
        if (k == "--model-a") model_a = v;
        else if (k == "--model-b") model_b = v;
        else if (k == "--model-path") model_path = v;
        else if (k == "--step") step = std::atoi(v);
        else if (k == "--opening-book") opening_book = v;
        else if (k == "--rounds") rounds = std::atoi(v);
        else if (k == "--board") cfg.board = std::atoi(v);
        else if (k == "--half-komi") cfg.half_komi = std::atoi(v);
        else if (k == "--games") cfg.games = std::atoi(v);
        else if (k == "--device") cfg.device = std::atoi(v);
        else if (k == "--max-moves") cfg.max_moves = std::atoi(v);
        else if (k == "--sampled-actions") cfg.sampled_actions = std::atoi(v);
        else if (k == "--budget") cfg.budget = (unsigned)std::atoi(v);
        else if (k == "--arena-slots") cfg.arena_slots = (unsigned)std::atoi(v);
        else if (k == "--seed") cfg.seed = std::strtoull(v, nullptr, 10);
        else {
            std::fprintf(stderr, "unknown flag %s\n", k.c_str());
            return 2;
        }
    }
    try {
        const int G = cfg.games;
        BatchedMCTS first(cfg.board, cfg.half_komi, G, cfg.device, 0, cfg.arena_slots);
        BatchedMCTS second(cfg.board, cfg.half_komi, G, cfg.device, 0, cfg.arena_slots);
        std::vector<tz_state_t> book;
        if (!opening_book.empty()) {  // one TPS per line (main.rs:146-160)
            std::ifstream f(opening_book);
            if (!f) throw std::runtime_error("Path to opening book should be valid");
            std::string line;
            while (std::getline(f, line)) {
                if (line.empty()) continue;
                tz_state_t g;
                if (!parse_tps(line, cfg.board, &g)) throw std::runtime_error("Opening book should be valid TPS, one per line");
                book.push_back(g);
            }
            if ((int)book.size() < G)
                throw std::runtime_error("There should be enough games in the opening book to form a unique batch");
        }
        for (int round = 0; round < rounds; round++) {
            std::string path_a = model_a, path_b = model_b;
            if (!model_path.empty()) {  // a random match-up among the directory's models (main.rs:163-185)
                const std::vector<std::string> paths = model_files(model_path, step);
                if (paths.size() < 2) {
                    std::fprintf(stderr, "Too few models.\n");
                    return 1;
                }
                const size_t a = mix64(cfg.seed * 0x9e3779b97f4a7c15ULL + (uint64_t)round) % paths.size();
                const size_t b = (a + 1 + mix64(cfg.seed + 0x632be59bd9b4e019ULL * (uint64_t)(round + 1)) % (paths.size() - 1)) % paths.size();
                path_a = paths[a];
                path_b = paths[b];
            }
            try {
                use_model(first, path_a);
                use_model(second, path_b);
            } catch (const std::exception& e) {  // `Cannot load {path}` -> next match-up (main.rs:177-185)
                std::fprintf(stderr, "Cannot load %s / %s: %s\n", path_a.c_str(), path_b.c_str(), e.what());
                continue;
            }
            // starting positions (main.rs:189-204)
            std::vector<tz_state_t> games;
            if (!book.empty()) {
                std::unordered_set<size_t> taken;
                for (int g = 0; g < G; g++)
                    for (uint64_t attempt = 0;; attempt++) {
                        const size_t idx = mix64(cfg.seed * 0x9e3779b97f4a7c15ULL + (uint64_t)round * 1000003ULL + (uint64_t)g +
                                                 attempt * 0x632be59bd9b4e019ULL) % book.size();
                        if (taken.insert(idx).second) {
                            games.push_back(book[idx]);
                            break;
                        }
                    }
            } else {
                const uint64_t seed = cfg.seed + 7919ULL * (uint64_t)round;
                first.new_openings(seed);
                first.random_steps(2, seed);
                std::vector<uint8_t> third(G);  // steps = rng.random_range(2..=3)
                for (int g = 0; g < G; g++) third[g] = (uint8_t)(mix64(seed * 0x9e3779b97f4a7c15ULL + (uint64_t)g) & 1);
                first.random_steps(1, seed + 1, third);
                games = first.envs();
            }
            const std::string name_a = path_a.empty() ? "synthetic-a" : base_name(path_a);
            const std::string name_b = path_b.empty() ? "synthetic-b" : base_name(path_b);
            const Evaluation a_as_white = compete(first, second, games, cfg, 2 * (uint64_t)round);
            std::printf("%s vs. %s: %s\n", name_a.c_str(), name_b.c_str(), a_as_white.to_string().c_str());
            const Evaluation b_as_white = compete(second, first, games, cfg, 2 * (uint64_t)round + 1);
            std::printf("%s vs. %s: %s\n", name_b.c_str(), name_a.c_str(), b_as_white.to_string().c_str());
            std::fflush(stdout);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "evaluation: %s\n", e.what());
        return 1;
    }
    return 0;
}
