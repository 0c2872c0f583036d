// Workload modules: device state whose layout is specialised beyond the
// generic SoA table.
#pragma once
#include "engine.hpp"
#include "kernels.hpp"

namespace incerto
{

struct StepOp
{
    const char* name = "";
    // advance `nsteps` steps starting at StepNumber == first_step
    std::function<int32_t(Sim&, uint64_t first_step, uint32_t nsteps)> fn;
    bool multi_step_ok = false; // entity-local: may apply several steps per launch
    bool whole_run_ok = false;  // fn may be handed every step of a run at once (it samples its own fused series)
    // host-side preparation (graph capture ...) done before the run's timing events are recorded
    std::function<int32_t(Sim&, uint64_t nsteps)> prepare;
    bool bumps_device_step = false;
};

struct ForestModule
{
    bool active = false;
    int table = -1;
    int comp_pos = -1, comp_cell = -1;
    int32_t W = 0, H = 0, x_begin = 0, x_end = 0;
    uint32_t pitch = 0;
    size_t slab_bytes = 0;          // (slab_w + 2) * pitch
    uint8_t* d_alloc[2] = {nullptr, nullptr}; // with 256 B lead/tail padding
    uint8_t* d_slab[2] = {nullptr, nullptr};  // column 0 of each buffer
    int cur = 0;
    uint32_t n_overlay = 0;
    std::vector<int16_t> h_ov_pos;
    std::vector<uint8_t> h_ov_owner;
    int16_t* d_ov_pos = nullptr;
    uint8_t* d_ov[2] = {nullptr, nullptr};
    uint8_t* d_ov_lo = nullptr; // overlay states received from the neighbours (row partition)
    uint8_t* d_ov_hi = nullptr;
    uint8_t* d_ov_owner = nullptr;
    uint32_t* d_near_bits = nullptr; // one bit per CTA tile with an overlay entity within one cell
    bool do_spread = false, do_progress = false, do_regrow = false;
    double p_spread = 0, p_regrow = 0;
    uint32_t burn_duration = 3;
    uint64_t* d_stats_run = nullptr; // running (healthy, burning, empty) totals kept by the step kernel
    uint64_t* d_stats = nullptr; // 5 u64 of the last step
    int series = -1;             // index of the fused FireStats series, -1 if none
    bool stats_valid = false;
    uint64_t unpublished_step = 0;   // last step whose FireStats record is not written yet (0: none); forest_finalize
    uint64_t d_step_mirror = ~0ull;  // what the device StepNumber holds (only captured graphs of steps read it)
    // captured step program: `graph_steps` consecutive steps (even, so the ping-pong buffers line up)
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_steps = 0;
    int graph_cur = 0;
    const void* graph_series_values = nullptr;
    uint64_t graph_series_capacity = 0;
    // halo transport over peer memory (row partition): the neighbours' slabs, overlay inboxes and mailboxes mapped
    // through CUDA IPC; `peer` stays false when that is not possible and NCCL send/recv is used instead
    bool peer = false;
    bool fused = true;               // the step kernel itself waits / pushes / stamps (else: separate kernels)
    uint8_t* peer_alloc_lo[2] = {nullptr, nullptr}; // lower neighbour's d_alloc[0/1]
    uint8_t* peer_alloc_hi[2] = {nullptr, nullptr};
    uint8_t* peer_ov_lo = nullptr;   // lower neighbour's d_ov_hi (my overlay states are its upper neighbour's)
    uint8_t* peer_ov_hi = nullptr;   // upper neighbour's d_ov_lo
    uint8_t* peer_ovs_lo[2] = {nullptr, nullptr}; // the neighbours' own overlay state arrays d_ov[0/1] (pulled by the fused transport)
    uint8_t* peer_ovs_hi[2] = {nullptr, nullptr};
    uint32_t* peer_mail_lo = nullptr; // lower neighbour's mailbox
    uint32_t* peer_mail_hi = nullptr;
    uint32_t* d_mail = nullptr;      // [0] steps pushed by the lower neighbour, [1] by the upper one
    uint32_t peer_slab_w_lo = 0;     // lower neighbour's slab width (to find its upper halo column)
    bool graph_pdl = false;          // the captured steps are programmatic dependents of each other
    uint32_t pushes = 0;             // steps whose boundary I pushed so far
    uint32_t rendezvous = 0;         // epochs of forest_cold_rendezvous
    void* ipc_opened[12] = {};
};

struct PandemicModule
{
    bool active = false;
    int table = -1;
    int comp_person = -1, comp_position = -1, comp_infected = -1;
    uint32_t grid_size = 0;
    bool do_move = false, do_infect = false, do_die = false, do_recover = false;
    double p_infect = 0, p_die = 0, p_recover = 0;
    uint32_t* d_cell_bits = nullptr;
    uint32_t* d_cell_head = nullptr;
    uint32_t* d_next = nullptr;
    uint64_t bits_vec4 = 0;    // 16-byte units of ONE of the two cell bitmaps in d_cell_bits (step parity selects)
    uint64_t head_slots = 0;   // entries of d_cell_head
    uint32_t head_shift = 0;   // 0: exact (one slot per cell), else the hash shift
};

struct AirModule
{
    bool active = false;
    int table = -1;
    int comp_pos = -1, comp_target = -1;
    int32_t size = 0, height = 0;
    bool do_check = false, do_move = false;
    double p_move = 0.7;
    int grid = -1;                        // SpatialGrid3D<Aircraft> the conflict check reads
    unsigned long long* d_slots = nullptr; // per-step occupancy hash (kernels_air.cu)
    uint32_t n_slots = 0, slot_bits = 0;
    uint32_t* d_bits = nullptr;           // hashed occupancy bitmap (pre-filter), 2^filter_bits bits
    uint32_t* d_bits_next = nullptr;      // the same for the next step, filled by the step kernel
    bool bits_ready = false;              // d_bits already describes the current positions
    uint32_t filter_bits = 0;
    uint8_t* d_cand = nullptr;            // per row: neighbour cells that showed a set bit
    uint64_t* d_acc = nullptr;            // [0] conflicts of the last step (pairs counted twice), [1] aircraft indexed
    // domain decomposition by x (incerto_builder_set_row_partition on an air_traffic world): see module_air.cu
    bool part = false;
    int32_t x_lo = 0, x_hi = 0;           // owned slab in grid coordinates
    uint32_t mail_cap = 0;                // records per mailbox list
    uint8_t* d_mail = nullptr;            // four AirMailbox blobs: out_lo, out_hi, in_lo, in_hi
    uint32_t* d_row_of_uid = nullptr;     // inverse of the spawn-time row sort (nullptr: row == uid)
};

struct TradersModule
{
    bool active = false;
    int table_traders = -1, table_stocks = -1;
    incerto_traders_params p{};
    bool do_buy = false, do_sell = false, do_net_worth = false, do_price = false;
    uint32_t num_stocks_cap = 0; // portfolio capacity in stocks (fixed at build)
};

struct SpatialModule
{
    bool active = false;
    int table = -1, grid = -1;
    incerto_spatial_params p{};
    uint32_t systems = 0;          // bit k: system kind 40 + k registered
    uint8_t* d_aux = nullptr;      // per person: ContactHistory length | pending grid op
    uint32_t* d_count = nullptr;   // persons per cell (PreUpdate snapshot)
    uint32_t* d_src_bits = nullptr;
    uint32_t* d_src_head = nullptr;
    uint32_t* d_src_next = nullptr;
    unsigned long long* d_mark = nullptr;
    uint8_t* d_marked = nullptr;   // 1 bit per person: its marks are in d_mark
    uint64_t* d_stats_acc = nullptr; // PandemicStats counts of the state after the last step (fused into the interact kernel)
    bool stats_valid = false;
};

struct CoinModule
{
    bool active = false;
    int table = -1;
    int comp_heads = -1, comp_tosses = -1;
    double odds = 0.5;
};

struct Sim::Modules
{
    ForestModule forest;
    CoinModule coin;
    PandemicModule pandemic;
    AirModule air;
    TradersModule traders;
    SpatialModule spatial;
    std::vector<StepOp> program;
    uint32_t* d_step = nullptr; // device mirror of StepNumber
};

// engine.cu
struct GridDesc;
int32_t table_sort_by_cell(Sim& s, int table, int pos_comp, const GridDesc& g, const std::vector<int>& zero_comps);

// module_forest.cu
int32_t forest_spawn_device(Sim& s, const HostSpawn& sp);
int32_t forest_bind(Sim& s);
int32_t forest_reset_after_spawn(Sim& s);
int32_t forest_step(Sim& s, uint64_t first_step, uint32_t nsteps);
int32_t forest_prepare(Sim& s, uint64_t nsteps);
int32_t forest_read_cells(Sim& s, uint8_t* out_rows); // W*H + overlays, uid order (this rank's slab only in partition mode)
int32_t forest_stats_now(Sim& s, uint64_t out5[5]);
void forest_free(Sim& s);
int32_t forest_halo_exchange(Sim& s);
int32_t forest_finalize(Sim& s);
int32_t air_mail_export(Sim& s, uint32_t side, void* out, size_t capacity, size_t* n_out);
int32_t air_mail_import(Sim& s, uint32_t side, const void* in, size_t n);
int32_t forest_peer_setup(Sim& s);
int32_t forest_peer_quiesce(Sim& s);
int32_t forest_cold_rendezvous(Sim& s); // run_cold: line the neighbours of a row-split grid up before a timed step
void forest_peer_free(Sim& s);

// module_pandemic.cu
int32_t pandemic_bind(Sim& s);
int32_t pandemic_step(Sim& s, uint64_t first_step, uint32_t nsteps);
int32_t pandemic_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset);
void pandemic_free(Sim& s);
int32_t pandemic_reset(Sim& s);

// module_air.cu
int32_t air_bind(Sim& s);
int32_t air_step(Sim& s, uint64_t first_step, uint32_t nsteps);
int32_t air_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset);
void air_free(Sim& s);
int32_t air_reset(Sim& s);

// module_traders.cu
int32_t traders_layout(Sim& s, const std::vector<uint64_t>& rows);
int32_t traders_bind(Sim& s);
int32_t traders_step(Sim& s, uint64_t first_step, uint32_t nsteps);
int32_t traders_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset);
int32_t traders_read_portfolio(Sim& s, int table, uint8_t* out, size_t out_size, size_t* written);

// module_spatial.cu
int32_t spatial_layout(Sim& s);
int32_t spatial_bind(Sim& s);
int32_t spatial_reset(Sim& s);
int32_t spatial_step(Sim& s, uint64_t first_step, uint32_t nsteps);
int32_t spatial_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset);
int32_t spatial_read_history(Sim& s, int table, uint8_t* out, size_t out_size, size_t* written);
int32_t spatial_stats_now(Sim& s, uint64_t initial_population, uint8_t* d_raw);
void spatial_free(Sim& s);

// comm.cpp
int32_t comm_unique_id(uint8_t out[128]);
int32_t comm_attach(Sim& s, const uint8_t id[128], uint32_t rank, uint32_t world);
void comm_free(Sim& s);
int32_t comm_allreduce(Sim& s, void* d_buf, uint32_t n, int is_f64, uint32_t op);
int32_t comm_allgather_bytes(Sim& s, const void* d_send, void* d_recv, size_t bytes);
// n 64-bit host values (u64 or f64) reduced over the communicator in place; op 0 sum, 1 min, 2 max
int32_t comm_allreduce_host(Sim& S, void* values, uint32_t n, int is_f64, uint32_t op);
// send `bytes_lo` from d_send_lo to rank-1 and `bytes_hi` from d_send_hi to rank+1, receive likewise
int32_t comm_halo(Sim& s, const void* const* d_send_lo, void* const* d_recv_lo, const void* const* d_send_hi,
                  void* const* d_recv_hi, const size_t* bytes, uint32_t n_msgs);

} // namespace incerto
