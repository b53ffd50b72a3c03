// format_check.cpp -- CPU-only test driver for the host-side formatting / parsing of
// include/takzero_b200.hpp (no GPU call): reads commands from stdin, one per line, prints results.
//   tps <n> <tps...>      -> round trip parse_tps / tps
//   move <text>           -> round trip parse_move / move_to_string, prints "<u16> <text>"
//   f32 <hex bits>        -> format_f32 of the float with these bits
//   replay <n> <line...>  -> Replay::parse then to_string (no result suffix)
//   target <n> <line...>  -> Target::parse then to_string, then " | <value bits> <ube bits> <policy bits...>"
//   eval <tag> <ply> <negations> -> Eval walk-back: negate k times, print f32 bits
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "../include/takzero_b200.hpp"

using namespace takzero;

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        std::string cmd;
        in >> cmd;
        if (cmd == "tps") {
            int n;
            in >> n;
            std::string rest;
            std::getline(in, rest);
            rest = rest.substr(rest.find_first_not_of(' '));
            tz_state_t g;
            if (!parse_tps(rest, n, &g)) {
                std::cout << "ERR" << std::endl;
                continue;
            }
            std::cout << tps(g, n) << " | " << (int)g.stones[0] << ' ' << (int)g.stones[1] << ' ' << (int)g.caps[0] << ' '
                      << (int)g.caps[1] << ' ' << g.ply << std::endl;
        } else if (cmd == "move") {
            std::string t;
            in >> t;
            Move m;
            if (!parse_move(t, &m)) std::cout << "ERR" << std::endl;
            else std::cout << m << ' ' << move_to_string(m) << std::endl;
        } else if (cmd == "f32") {
            unsigned bits;
            in >> std::hex >> bits;
            float v;
            std::memcpy(&v, &bits, 4);
            std::cout << format_f32(v) << std::endl;
        } else if (cmd == "replay") {
            int n;
            in >> n;
            std::string rest;
            std::getline(in, rest);
            Replay r;
            if (!Replay::parse(rest, n, &r)) std::cout << "ERR" << std::endl;
            else std::cout << r.to_string(n, 0);
        } else if (cmd == "target") {
            int n;
            in >> n;
            std::string rest;
            std::getline(in, rest);
            rest = rest.substr(rest.find_first_not_of(' '));
            Target t;
            if (!Target::parse(rest, n, &t)) {
                std::cout << "ERR" << std::endl;
                continue;
            }
            std::string text = t.to_string(n);
            text.pop_back();
            auto bits = [](float v) {
                unsigned b;
                std::memcpy(&b, &v, 4);
                return b;
            };
            std::printf("%s |%08x %08x", text.c_str(), bits(t.value), bits(t.ube));
            for (auto& mp : t.policy) std::printf(" %u:%08x", (unsigned)mp.first, bits(mp.second));
            std::printf("\n");
            std::fflush(stdout);
        } else if (cmd == "eval") {
            unsigned tag, ply;
            int k;
            in >> tag >> ply >> k;
            Eval e;
            e.tag = tag;
            e.ply = ply;
            for (int i = 0; i < k; i++) e = e.negate();
            const float v = e.to_f32();
            unsigned bits;
            std::memcpy(&bits, &v, 4);
            std::printf("%u %u %08x\n", e.tag, e.ply, bits);
        }
    }
    return 0;
}
