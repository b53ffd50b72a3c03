// reanalyze.cpp -- the reference's `reanalyze` binary (reanalyze/src/main.rs:60-244) as a C++ host over
// libtakzero_b200.so.  Every iteration: wait while the trainer's reanalyze buffer is full (`buffer_lengths.txt`,
// main.rs:78-91), reload `model_latest.ot` (main.rs:93), read the replays appended to `replays.txt` since the
// last iteration from the remembered byte offset and expand them into positions (main.rs:262-285; with
// --exploration-positions also `replays-exploration.txt` into a bounded buffer, main.rs:117-133,148-153), sample a
// batch of distinct positions as fresh roots, search them with Gumbel sequential halving (beta = 0) and append
// `targets-reanalyze.txt` with
//   value  = root evaluation if solved, else -evaluation of the selected child   (main.rs:184-195)
//   policy = improved_policy(most_visited_count())                               (main.rs:196-202)
//   ube    = ube_target(0.25)                                                     (main.rs:203)
// Differences at the process boundary: constants are flags (MIN_POSITIONS defaults to the batch size instead of
// 128000); positions are sampled without replacement by a counter-based hash of (seed, batch, slot) instead of
// `rand`'s `sample` (whose stream is not pinned); the model is reloaded only when the file changed; with too few
// positions the program exits instead of sleeping 60 s unless --wait is given; it stops after --batches iterations.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
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

// `count` distinct indices below `n` for batch `b`: slot g takes the first value of its hash sequence that no
// earlier slot took
static std::vector<size_t> sample_distinct(uint64_t seed, uint64_t b, int first_slot, int count, size_t n) {
    std::vector<size_t> out;
    std::unordered_set<size_t> taken;
    for (int g = first_slot; g < first_slot + count; g++)
        for (uint64_t attempt = 0;; attempt++) {
            const size_t idx = mix64(seed * 0x9e3779b97f4a7c15ULL + b * 1000003ULL + (uint64_t)g + attempt * 0x632be59bd9b4e019ULL) % n;
            if (taken.insert(idx).second) {
                out.push_back(idx);
                break;
            }
        }
    return out;
}

// fill_buffer_with_positions_from_replays (main.rs:262-285): parse the lines after byte offset *seek, append
// Replay::states of each; unparsable lines are skipped like `parse().ok()` does
static void fill_buffer(BatchedMCTS& mcts, std::vector<tz_state_t>& buffer, std::streamoff* seek, const std::string& path,
                        int board) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return;
    f.seekg(*seek);
    std::string line;
    std::vector<Replay> replays;
    while (std::getline(f, line)) {
        if (f.eof() && !line.empty()) break;  // a line still being written: read it next time
        *seek += (std::streamoff)line.size() + 1;
        Replay r;
        if (Replay::parse(line, board, &r)) replays.push_back(std::move(r));
    }
    const std::vector<tz_state_t> st = mcts.replay_states(replays);
    buffer.insert(buffer.end(), st.begin(), st.end());
}

static const size_t MAX_REANALYZE_BUFFER_LEN = 32000;  // main.rs:41

int main(int argc, char** argv) {
    std::string directory = ".", weights;
    int board = 6, half_komi = 4, games = 128, device = 0, batches = 1, sampled_actions = 64;
    int exploration_positions = 0;                      // EXPLORATION_POSITIONS_IN_BATCH (main.rs:46)
    size_t max_exploration_buffer = 0, min_positions = 0;  // MAX_EXPLORATION_BUFFER_SIZE, MIN_POSITIONS
    bool wait = false;
    unsigned budget = 768, arena_slots = 0;
    unsigned long long seed = 1;
    float ube_beta = 0.25f;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        if (k == "--wait") {
            wait = true;
            continue;
        }
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
        else if (k == "--min-positions") min_positions = std::strtoull(v, nullptr, 10);
        else if (k == "--exploration-positions") exploration_positions = std::atoi(v);
        else if (k == "--max-exploration-buffer") max_exploration_buffer = std::strtoull(v, nullptr, 10);
        else {
            std::fprintf(stderr, "unknown flag %s\n", k.c_str());
            return 2;
        }
    }
    if (min_positions < (size_t)games) min_positions = (size_t)games;
    if (exploration_positions < 0 || exploration_positions > games) {
        std::fprintf(stderr, "--exploration-positions must be within the batch\n");
        return 2;
    }
    try {
        BatchedMCTS mcts(board, half_komi, games, device, 0, arena_slots);
        for (const char* name : {"/model_latest.ot", "/model_latest.tzw"})
            if (weights.empty() && mtime_ns(directory + name) >= 0) weights = directory + name;
        long long model_stamp = -1;
        std::vector<tz_state_t> position_buffer, exploration_buffer;
        std::streamoff replays_seek = 0, exploration_replays_seek = 0;
        const int stride = mcts.move_stride();
        for (int b = 0; b < batches;) {
            for (;;) {  // main.rs:78-91
                const BufferLengths lengths = read_buffer_lengths(directory);
                if (lengths.status == -1 || (lengths.status == 0 && lengths.reanalyze <= MAX_REANALYZE_BUFFER_LEN)) break;
                if (lengths.status == -2) std::fprintf(stderr, "Could not read buffer lengths: wrong checksum or missing component\n");
                std::this_thread::sleep_for(std::chrono::seconds(1));
            }
            if (!weights.empty()) {  // Net::load (main.rs:93), skipped while the file is unchanged
                const long long stamp = mtime_ns(weights);
                if (stamp != model_stamp) try {
                        mcts.load_model(weights);
                        mcts.set_agent(TZ_AGENT_NETWORK);
                        model_stamp = stamp;
                    } catch (const std::exception& e) {
                        if (model_stamp < 0) throw;
                        std::fprintf(stderr, "Cannot load model: %s, keeping the previous one.\n", e.what());
                    }
            }
            fill_buffer(mcts, position_buffer, &replays_seek, directory + "/replays.txt", board);
            if (exploration_positions > 0) {
                fill_buffer(mcts, exploration_buffer, &exploration_replays_seek, directory + "/replays-exploration.txt", board);
                if (exploration_buffer.size() > max_exploration_buffer)
                    exploration_buffer.erase(exploration_buffer.begin(),
                                             exploration_buffer.end() - (std::ptrdiff_t)max_exploration_buffer);
            }
            if (position_buffer.size() < min_positions) {
                std::fprintf(stderr, "reanalyze: not enough positions yet (%zu)\n", position_buffer.size());
                if (!wait) return 1;
                std::this_thread::sleep_for(std::chrono::seconds(60));
                continue;
            }
            // sample a batch (main.rs:146-165): exploration positions first, the rest from the position buffer
            std::vector<tz_state_t> batch;
            const int from_exploration = std::min<size_t>((size_t)exploration_positions, exploration_buffer.size());
            for (size_t idx : sample_distinct(seed ^ 0x5851f42d4c957f2dULL, (uint64_t)b, 0, from_exploration, exploration_buffer.size()))
                batch.push_back(exploration_buffer[idx]);
            for (size_t idx : sample_distinct(seed, (uint64_t)b, 0, games - from_exploration, position_buffer.size()))
                batch.push_back(position_buffer[idx]);
            mcts.set_positions(batch);  // *node = Node::default(); *env = replay_env
            // search and targets on the device (tz_reanalyze_batch): improved policy at most_visited_count() visitations,
            // UBE target, and the value rule of main.rs:184-195 (root evaluation if known, else the negated evaluation
            // of the selected child) -- only the targets themselves come back
            mcts.reanalyze_batch(nullptr, sampled_actions, budget, ube_beta, seed + (uint64_t)b);
            const BatchedMCTS::ReanalyzeTargets rt = mcts.reanalyze_read();
            std::string contents;
            for (int g = 0; g < games; g++) {
                Target t;
                t.env = batch[g];
                for (int i = 0; i < rt.n[g]; i++)
                    t.policy.emplace_back(rt.moves[(size_t)g * stride + i], rt.policy[(size_t)g * stride + i]);
                t.value = rt.value[g];
                t.ube = rt.ube[g];
                contents += t.to_string(board);
            }
            if (!append_file(directory + "/targets-reanalyze.txt", contents))
                std::fprintf(stderr, "Could not save targets to file, so here they are instead:\n%s", contents.c_str());
            b++;
        }
        std::printf("reanalyze: %d batches of %d positions out of %zu\n", batches, games, position_buffer.size());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "reanalyze: %s\n", e.what());
        return 1;
    }
    return 0;
}
