//! Host-side drop-in for the reference's batched search on top of libtakzero_b200.so.
//!
//! UNVERIFIED: the image this repository is built in has no `rustc` / `cargo`, so this crate has never been
//! compiled.  `sys.rs` is generated from `include/takzero_b200.h` (tools/gen_rust_bindings.py) and checked against
//! the header and the shared library's exports by `tests/test_rust_bindings.py`; this file is the hand-written
//! wrapper a maintainer of ViliamVadocz/takzero would start from.  The tested host layers are the C++ one
//! (`include/takzero_b200.hpp`, `host/*.cpp`) and the ctypes one (`takzero_b200/capi.py`).
//!
//! Mapping (reference item -> here):
//! * `BatchedMCTS<BATCH_SIZE, E>` (takzero/src/search/node/batched.rs:24-409) -> [`BatchedMcts`]; the batch size is
//!   a run-time value and the trees live in GPU memory, so `nodes_and_envs()` becomes the read-backs
//!   [`BatchedMcts::root_children`] / [`BatchedMcts::root_stats`] / [`BatchedMcts::envs`].
//! * `Net::load(path, device)` (network/mod.rs:20-27, net6_simhash.rs:164-181) -> [`BatchedMcts::load_model`],
//!   which reads `model_latest.ot` and `bitvec.bin` itself (no tch on this path).
//! * `Environment` (search/env.rs:11-25) -> [`BatchedMcts::legal_moves`], [`BatchedMcts::apply`],
//!   [`BatchedMcts::terminal`] on [`sys::tz_state_t`]; implement [`ToState`] for `fast_tak::Game<N, HALF_KOMI>`.
//! * `Eval` (search/eval.rs:8-13) <- `(eval_tag, eval_bits)`: 0 `Value(f32::from_bits)`, 1 `Win(ply)`, 2 `Loss(ply)`,
//!   3 `Draw(ply)`.
//! Every call returns `Err(message)` where the reference would panic (`env.rs:44`, `batched.rs:215-220`,
//! `net6_simhash.rs:304`); `unwrap()` keeps the reference's behaviour.
pub mod sys;

use std::ffi::{CStr, CString};
use std::ptr;

pub type Move = sys::tz_move_t;
pub type State = sys::tz_state_t;

/// takparse `Move` -> the library's 2-byte move: bits 0..2 file, 3..5 rank, 6..7 piece (flat 0, wall 1, cap 2) or
/// direction (`+` 0, `-` 1, `<` 2, `>` 3), bits 8..15 `Pattern::mask()` (0 for placements).
pub fn pack_move(col: u8, row: u8, kind_or_dir: u8, pattern_mask: u8) -> Move {
    (col as u16) | ((row as u16) << 3) | ((kind_or_dir as u16) << 6) | ((pattern_mask as u16) << 8)
}

/// Implement for `fast_tak::Game<N, HALF_KOMI>`: square `row * N + col`; `height = stack.size()`, `top` = piece of
/// the top stone (flat 0, wall 1, cap 2), bit `i` of `stack` = colour (1 = black) of the `i`-th stone from the
/// bottom; plus `to_move`, reserves, `ply`, `reversible_plies` (the fields the reference reads in repr.rs:177-223).
pub trait ToState {
    fn to_state(&self) -> State;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::tz_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: i32) -> Result<i32, String> {
    if rc < 0 { Err(format!("{} (code {rc})", last_error())) } else { Ok(rc) }
}

pub struct RootChildren {
    pub stride: usize,
    pub n: Vec<i32>,
    pub moves: Vec<Move>,
    pub visits: Vec<u32>,
    pub eval_tag: Vec<u32>,
    pub eval_bits: Vec<u32>,
    pub logit: Vec<f32>,
    pub prob: Vec<f32>,
    pub std_dev: Vec<f32>,
}

/// `BatchedMCTS`: all games of one GPU advance in lock-step; game `g` has the global id `game_base + g`
/// (RNG streams are keyed by it, so sharded runs equal unsharded ones).
pub struct BatchedMcts {
    h: *mut sys::tz_handle,
    games: usize,
    stride: usize,
}

impl BatchedMcts {
    /// `BatchedMCTS::new` (batched.rs:33-47); openings come from [`Self::new_openings`].
    pub fn new(board_n: i32, half_komi: i32, n_games: i32, device: i32, game_base: i32) -> Result<Self, String> {
        let cfg = sys::tz_config_t { board_n, half_komi, n_games, device, game_base, reversible_limit: 0,
                                     move_stride: 0, arena_slots: 0, tree_batch: 0 };
        let mut h = ptr::null_mut();
        check(unsafe { sys::tz_create(&cfg, &mut h) })?;
        let mut stride = 0;
        check(unsafe { sys::tz_info(h, &mut stride, ptr::null_mut(), ptr::null_mut(), ptr::null_mut()) })?;
        Ok(Self { h, games: n_games as usize, stride: stride as usize })
    }
    /// `Net::load` + `Agent for Net`: later searches evaluate leaves with the device network.
    pub fn load_model(&mut self, path: &str) -> Result<(), String> {
        let c = CString::new(path).map_err(|e| e.to_string())?;
        check(unsafe { sys::tz_load_model(self.h, c.as_ptr()) })?;
        check(unsafe { sys::tz_set_agent(self.h, sys::TZ_AGENT_NETWORK as i32, None, ptr::null_mut()) }).map(|_| ())
    }
    /// One process per GPU (csrc/comm.cu): rank 0 calls [`BatchedMcts::comm_unique_id`] and hands the 128 bytes to the
    /// others (a file in the run directory will do), then every rank calls `comm_init`.
    pub fn comm_unique_id() -> Result<[u8; 128], String> {
        let mut id = [0u8; 128];
        check(unsafe { sys::tz_comm_unique_id(id.as_mut_ptr() as *mut _) })?;
        Ok(id)
    }
    pub fn comm_init(&mut self, id: &[u8; 128], nranks: i32, rank: i32) -> Result<(), String> {
        check(unsafe { sys::tz_comm_init(self.h, id.as_ptr() as *const _, nranks, rank) }).map(|_| ())
    }
    /// The `Net::load` before every move (selfplay/src/main.rs:107) as ONE collective over all ranks: the root passes
    /// the model's tensors (e.g. read with `tz_read_model_file`), the others `None`; every rank swaps weight sets
    /// between two moves.  Without a communicator this is `Net::load` on one GPU.
    pub fn broadcast_weights(&mut self, tensors: Option<&[sys::tz_tensor_t]>, res_blocks: i32, root: i32) -> Result<(), String> {
        let (p, n) = match tensors { Some(t) => (t.as_ptr(), t.len() as i32), None => (ptr::null(), 0) };
        check(unsafe { sys::tz_broadcast_weights(self.h, p, n, res_blocks, root) }).map(|_| ())
    }
    /// `Net::update_counts` (net6_simhash.rs:236-241): the novelty set remembers these positions.
    pub fn update_counts(&mut self, states: &[State]) -> Result<(), String> {
        check(unsafe { sys::tz_update_counts(self.h, states.as_ptr(), states.len() as i32) }).map(|_| ())
    }
    /// The set as `Net::save` writes it next to the model (`bitvec.bin`, 2^29 bytes; net6_simhash.rs:152-170).
    pub fn novelty_set(&self) -> Result<Vec<u8>, String> {
        let mut out = vec![0u8; 1 << 29];
        check(unsafe { sys::tz_read_novelty_set(self.h, out.as_mut_ptr(), out.len()) })?;
        Ok(out)
    }
    /// Whole-job totals of per-rank counters (what `learn` adds up through `buffer_lengths.txt`).
    pub fn allreduce_sum(&mut self, values: &mut [u64]) -> Result<(), String> {
        check(unsafe { sys::tz_allreduce_sum(self.h, values.as_mut_ptr(), values.len() as i32) }).map(|_| ())
    }

    /// `Env::new_opening` for every game, drawn by the library from `seed`.
    pub fn new_openings(&mut self, seed: u64) -> Result<(), String> {
        check(unsafe { sys::tz_new_openings(self.h, ptr::null(), ptr::null(), ptr::null(), seed) }).map(|_| ())
    }
    /// `BatchedMCTS::from_envs` / `nodes_and_envs_mut` writes (reanalyze/src/main.rs:159-165): fresh roots.
    pub fn set_envs(&mut self, envs: &[State]) -> Result<(), String> {
        assert_eq!(envs.len(), self.games);
        check(unsafe { sys::tz_set_positions(self.h, envs.as_ptr(), ptr::null()) }).map(|_| ())
    }
    pub fn envs(&self) -> Result<Vec<State>, String> {
        let mut out = vec![unsafe { std::mem::zeroed::<State>() }; self.games];
        check(unsafe { sys::tz_get_positions(self.h, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// `BatchedMCTS::simulate` (batched.rs:63-128).
    pub fn simulate(&mut self, betas: &[f32]) -> Result<(), String> {
        assert_eq!(betas.len(), self.games);
        check(unsafe { sys::tz_simulate(self.h, betas.as_ptr()) }).map(|_| ())
    }
    /// `BatchedMCTS::gumbel_sequential_halving` (batched.rs:207-409).  `gumbel`: one Gumbel(0,1) draw per root child
    /// in child order, `[games][move_stride]` (pass the reference's `rng` draws for identical results), or `None`
    /// to let the library draw them from `seed`.
    pub fn gumbel_sequential_halving(&mut self, betas: &[f32], sampled_actions: usize, search_budget: u32,
                                     gumbel: Option<&[f32]>, seed: u64) -> Result<Vec<Move>, String> {
        assert_eq!(betas.len(), self.games);
        let mut out = vec![0 as Move; self.games];
        let g = gumbel.map_or(ptr::null(), |g| { assert_eq!(g.len(), self.games * self.stride); g.as_ptr() });
        check(unsafe { sys::tz_gumbel_sequential_halving(self.h, betas.as_ptr(), sampled_actions as i32, search_budget, g,
                                                         self.stride as i32, seed, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// `BatchedMCTS::step` (batched.rs:131-144): re-root at the chosen child (subtree kept), `env.step`.
    pub fn step(&mut self, actions: &[Move]) -> Result<(), String> {
        assert_eq!(actions.len(), self.games);
        check(unsafe { sys::tz_step(self.h, actions.as_ptr()) }).map(|_| ())
    }
    /// `BatchedMCTS::restart_terminal_envs` (batched.rs:185-203): per game 0 (running) or the `Terminal` of the
    /// finished game for the side to move (1 win, 2 loss, 3 draw); its replay is read with [`Self::finished_replay`].
    pub fn restart_terminal_envs(&mut self, seed: u64) -> Result<Vec<i32>, String> {
        let mut out = vec![0; self.games];
        check(unsafe { sys::tz_restart_terminal(self.h, ptr::null(), ptr::null(), seed, out.as_mut_ptr()) })?;
        Ok(out)
    }
    pub fn finished_replay(&self, game: usize) -> Result<(State, Vec<Move>), String> {
        let mut start = unsafe { std::mem::zeroed::<State>() };
        let mut moves = vec![0 as Move; sys::TZ_MAX_PLIES];
        let len = check(unsafe { sys::tz_finished_replay(self.h, game as i32, &mut start, moves.as_mut_ptr(),
                                                         sys::TZ_MAX_PLIES as i32) })?;
        moves.truncate(len as usize);
        Ok((start, moves))
    }
    /// `select_actions_in_selfplay` (batched.rs:165-183, node/mod.rs:170-207) with library randomness.
    pub fn select_actions_in_selfplay(&mut self, weighted_random_plies: u16, seed: u64) -> Result<Vec<Move>, String> {
        let mut out = vec![0 as Move; self.games];
        check(unsafe { sys::tz_select_selfplay(self.h, weighted_random_plies as i32, 32, 0.5, ptr::null(), seed,
                                               out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// `Node::improved_policy(visitations)` zipped with the children, and `Node::ube_target(beta)`, of every root.
    pub fn targets(&self, visitations: f32, beta: f32) -> Result<(Vec<f32>, Vec<f32>, Vec<i32>, Vec<Move>), String> {
        let cells = self.games * self.stride;
        let (mut policy, mut ube) = (vec![0f32; cells], vec![0f32; self.games]);
        let (mut n, mut moves) = (vec![0i32; self.games], vec![0 as Move; cells]);
        check(unsafe { sys::tz_targets(self.h, visitations, beta, self.stride as i32, policy.as_mut_ptr(), ube.as_mut_ptr(),
                                       n.as_mut_ptr(), moves.as_mut_ptr()) })?;
        Ok((policy, ube, n, moves))
    }
    /// Direct `Node` field reads of the callers (`node.children`, `.evaluation`, `.visit_count`, ...).
    pub fn root_children(&self) -> Result<RootChildren, String> {
        let cells = self.games * self.stride;
        let mut c = RootChildren { stride: self.stride, n: vec![0; self.games], moves: vec![0; cells], visits: vec![0; cells],
                                   eval_tag: vec![0; cells], eval_bits: vec![0; cells], logit: vec![0.0; cells],
                                   prob: vec![0.0; cells], std_dev: vec![0.0; cells] };
        check(unsafe { sys::tz_root_children(self.h, self.stride as i32, c.n.as_mut_ptr(), c.moves.as_mut_ptr(),
                                             c.visits.as_mut_ptr(), c.eval_tag.as_mut_ptr(), c.eval_bits.as_mut_ptr(),
                                             c.logit.as_mut_ptr(), c.prob.as_mut_ptr(), c.std_dev.as_mut_ptr()) })?;
        Ok(c)
    }
    pub fn root_stats(&self) -> Result<Vec<sys::tz_root_t>, String> {
        let mut out = vec![unsafe { std::mem::zeroed::<sys::tz_root_t>() }; self.games];
        check(unsafe { sys::tz_root_stats(self.h, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// `Environment::populate_actions` / `step` / `terminal` for host positions (parity hooks).
    pub fn legal_moves(&self, states: &[State]) -> Result<Vec<Vec<Move>>, String> {
        let mut moves = vec![0 as Move; states.len() * self.stride];
        let mut n = vec![0i32; states.len()];
        check(unsafe { sys::tz_legal_moves(self.h, states.as_ptr(), states.len() as i32, self.stride as i32,
                                           moves.as_mut_ptr(), n.as_mut_ptr()) })?;
        Ok((0..states.len()).map(|i| moves[i * self.stride..i * self.stride + n[i] as usize].to_vec()).collect())
    }
    pub fn apply(&self, states: &mut [State], moves: &[Move]) -> Result<(), String> {
        assert_eq!(states.len(), moves.len());
        let mut ok = vec![0i32; states.len()];
        check(unsafe { sys::tz_apply(self.h, states.as_mut_ptr(), moves.as_ptr(), states.len() as i32, ok.as_mut_ptr()) })?;
        if ok.iter().all(|&x| x != 0) { Ok(()) } else { Err("Action should be valid".into()) }
    }
    pub fn terminal(&self, states: &[State]) -> Result<Vec<i32>, String> {
        let mut out = vec![0i32; states.len()];
        check(unsafe { sys::tz_result(self.h, states.as_ptr(), states.len() as i32, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

impl Drop for BatchedMcts {
    fn drop(&mut self) {
        unsafe { sys::tz_destroy(self.h) }
    }
}
