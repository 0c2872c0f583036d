// Pandemic module: host side of kernels_pandemic.cu (examples/pandemic.rs).
#include <algorithm>
#include <cstdlib>

#include "kernels_grid.hpp"
#include "kernels_pandemic.hpp"
#include "modules.hpp"

namespace incerto
{

static int32_t pandemic_sort_rows(Sim& s);

static int32_t pandemic_check(Sim& s, PandemicModule& pm, const incerto_pandemic_params& p)
{
    auto bad = [&](int c) { return c < 0 || c >= static_cast<int>(s.comps.size()); };
    if (bad(p.person) || bad(p.position) || bad(p.infected)) return INCERTO_ERR_INVALID_ARGUMENT;
    const ComponentInfo& pc = s.comps[p.position];
    if (s.comps[p.person].dtype != INCERTO_MARKER || s.comps[p.infected].dtype != INCERTO_MARKER || pc.lanes != 2 ||
        (pc.dtype != INCERTO_U16 && pc.dtype != INCERTO_I16))
    {
        set_error("pandemic: Person/Infected must be markers and Position a 2-lane U16 component (device layout)");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (p.grid_size == 0 || p.grid_size > 32768)
    {
        set_error("pandemic: grid_size must be in 1..=32768");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (pm.comp_person >= 0 && (pm.comp_person != p.person || pm.comp_position != p.position ||
                                pm.comp_infected != p.infected || pm.grid_size != p.grid_size))
    {
        set_error("pandemic systems must agree on their components and GRID_SIZE");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    pm.comp_person = p.person;
    pm.comp_position = p.position;
    pm.comp_infected = p.infected;
    pm.grid_size = p.grid_size;
    s.comps[p.person].exists = true;
    s.comps[p.position].exists = true;
    s.comps[p.infected].exists = true; // the systems hold Query<.., With<Infected>>: the component is registered
    return INCERTO_OK;
}

int32_t pandemic_bind(Sim& s)
{
    PandemicModule& pm = s.mods->pandemic;
    for (auto& sys : s.systems)
    {
        if (sys.kind < INCERTO_SYS_PANDEMIC_MOVE || sys.kind > INCERTO_SYS_PANDEMIC_RECOVER) continue;
        if (sys.params.size() != sizeof(incerto_pandemic_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_pandemic_params p;
        std::memcpy(&p, sys.params.data(), sizeof(p));
        INC_TRY(pandemic_check(s, pm, p));
        switch (sys.kind)
        {
        case INCERTO_SYS_PANDEMIC_MOVE: pm.do_move = true; break;
        case INCERTO_SYS_PANDEMIC_INFECT:
            pm.do_infect = true;
            pm.p_infect = p.probability;
            break;
        case INCERTO_SYS_PANDEMIC_DIE:
            pm.do_die = true;
            pm.p_die = p.probability;
            break;
        default:
            pm.do_recover = true;
            pm.p_recover = p.probability;
            break;
        }
    }
    pm.table = table_for_components(s, {pm.comp_position});
    pm.active = true;
    if (pm.table < 0) return INCERTO_OK; // no person was spawned: the systems have nothing to iterate
    Table& t = s.tables[pm.table];
    if (t.n >= 0xffffffffull)
    {
        set_error("pandemic: at most 2^32-2 people per replica (32-bit row links)");
        return INCERTO_ERR_UNSUPPORTED;
    }
    if (pm.do_infect)
    {
        const uint64_t cells = static_cast<uint64_t>(pm.grid_size) * pm.grid_size;
        pm.bits_vec4 = cells / 128 + 1;
        INC_CUDA(dev_alloc(&pm.d_cell_bits, 2 * pm.bits_vec4 * 16));
        INC_CUDA(cudaMemsetAsync(pm.d_cell_bits, 0, 2 * pm.bits_vec4 * 16, s.stream));
        // list heads: one slot per cell while that stays small, else ~n/4 hashed slots (a power of two): with the
        // reference's density (0.4 people per cell, ~10 % infected) the exact table would be 10 x the population's own
        // bytes and every exchange a DRAM round trip; the hashed one stays in the L2
        uint64_t want = 4096;
        while (want < t.n / 4) want <<= 1;
        if (const char* e = std::getenv("INCERTO_B200_PANDEMIC_EXACT_HEADS"))
            if (e[0] == '1') want = cells;
        if (want >= cells || want > (1ull << 31))
        {
            pm.head_slots = cells;
            pm.head_shift = 0;
        }
        else
        {
            pm.head_slots = want;
            uint32_t lg = 0;
            while ((1ull << lg) < want) ++lg;
            pm.head_shift = 32 - lg;
        }
        INC_CUDA(dev_alloc(&pm.d_cell_head, pm.head_slots * 4));
        INC_CUDA(dev_alloc(&pm.d_next, std::max<uint64_t>(t.n, 1) * 4));
        INC_CUDA(cudaMemsetAsync(pm.d_cell_head, 0xff, pm.head_slots * 4, s.stream));
    }
    return pandemic_sort_rows(s);
}

// optional: rows in cell order, so that the healthy people's look-ups of "is anybody infected in my cell" walk the
// per-cell tables sector by sector (people move at most one cell per step)
static int32_t pandemic_sort_rows(Sim& s)
{
    PandemicModule& pm = s.mods->pandemic;
    if (pm.table < 0 || !pm.do_infect) return INCERTO_OK;
    // off by default: with uid != row the move kernel needs one Philox call per row instead of one per four rows,
    // which costs more (74 vs 60 us at 10 M people) than the interact kernel gains (56 vs 66 us)
    const char* e = std::getenv("INCERTO_B200_PANDEMIC_ROW_SORT");
    if (!e || e[0] != '1') return INCERTO_OK;
    GridDesc g;
    g.dim = 2;
    g.bmin[0] = g.bmin[1] = g.bmin[2] = 0;
    g.ext[0] = g.ext[1] = static_cast<int32_t>(pm.grid_size);
    g.ext[2] = 1;
    return table_sort_by_cell(s, pm.table, pm.comp_position, g, {});
}

// reset(): the step counter restarts, so heads stamped by the previous run would look current
int32_t pandemic_reset(Sim& s)
{
    PandemicModule& pm = s.mods->pandemic;
    if (pm.d_cell_head)
    {
        INC_CUDA(cudaMemsetAsync(pm.d_cell_head, 0xff, pm.head_slots * 4, s.stream));
        INC_CUDA(cudaMemsetAsync(pm.d_cell_bits, 0, 2 * pm.bits_vec4 * 16, s.stream));
    }
    return pandemic_sort_rows(s);
}

void pandemic_free(Sim& s)
{
    PandemicModule& pm = s.mods->pandemic;
    dev_free(pm.d_cell_bits);
    dev_free(pm.d_cell_head);
    dev_free(pm.d_next);
    pm.d_cell_bits = pm.d_cell_head = pm.d_next = nullptr;
}

static void pandemic_fill(Sim& s, PandemicModule& pm, PandemicArgs& a, Table& t, uint64_t step)
{
    std::memset(&a, 0, sizeof(a));
    a.pos = static_cast<uint32_t*>(t.col(pm.comp_position)->d);
    a.flags = t.d_flags;
    a.uid = t.d_uid;
    a.n = t.n;
    a.grid_size = pm.grid_size;
    a.step = static_cast<uint32_t>(step);
    a.bit_person = 1u << s.comps[pm.comp_person].marker_bit;
    a.bit_infected = 1u << s.comps[pm.comp_infected].marker_bit;
    a.do_move = pm.do_move;
    a.do_infect = pm.do_infect;
    a.do_die = pm.do_die;
    a.do_recover = pm.do_recover;
    const bool two = (!pm.do_die || pm.p_die * 256.0 <= 1.0) && (!pm.do_recover || pm.p_recover * 256.0 <= 1.0);
    a.fate_two_level = two ? 1u : 0u;
    a.thr_infect = bernoulli_threshold(pm.p_infect);
    a.thr_die = bernoulli_threshold(two ? pm.p_die * 256.0 : pm.p_die);
    a.thr_recover = bernoulli_threshold(two ? pm.p_recover * 256.0 : pm.p_recover);
    a.rk_move = make_round_keys(make_key(s.seed, s.replica, TAG_PANDEMIC_MOVE));
    a.rk_fate = make_round_keys(make_key(s.seed, s.replica, TAG_PANDEMIC_FATE));
    a.rk_infect = make_round_keys(make_key(s.seed, s.replica, TAG_PANDEMIC_INFECT));
    // two bitmaps: this step's (clean: zeroed by the previous step's interact kernel) and the next step's
    a.cell_bits = pm.d_cell_bits ? pm.d_cell_bits + (step & 1u) * pm.bits_vec4 * 4 : nullptr;
    a.cell_bits_clear = pm.d_cell_bits ? pm.d_cell_bits + ((step + 1) & 1u) * pm.bits_vec4 * 4 : nullptr;
    a.bits_vec4 = pm.bits_vec4;
    a.cell_head = pm.d_cell_head;
    a.head_shift = pm.head_shift;
    a.link_mask = t.n >= 0x00ffffffu ? 0xffffffffu : 0x00ffffffu;
    a.next_infected = pm.d_next;
}

int32_t pandemic_step(Sim& s, uint64_t first_step, uint32_t nsteps)
{
    PandemicModule& pm = s.mods->pandemic;
    if (pm.table < 0) return INCERTO_OK;
    Table& t = s.tables[pm.table];
    if (!t.n) return INCERTO_OK;
    for (uint32_t i = 0; i < nsteps; ++i)
    {
        const uint64_t step = first_step + i;
        PandemicArgs a;
        pandemic_fill(s, pm, a, t, step);
        if (pm.do_infect)
        {
            // the heads carry the step's stamp (8 bits) next to a 24-bit row: cleared when the stamp wraps; a population
            // too large for 24-bit rows uses the whole word for the row and clears the heads before every step
            if ((step & 0xffu) == 0 || t.n >= 0x00ffffffu) INC_CUDA(cudaMemsetAsync(pm.d_cell_head, 0xff, pm.head_slots * 4, s.stream));
        }
        if (pm.do_move || pm.do_infect)
        {
            // position r+w 8 B + flags r 1 B per person (+ scratch for the infected)
            KTimer kt(s, "pandemic_move", 9.0 * t.n);
            launch_pandemic_move(a, s.stream);
        }
        if (pm.do_infect || pm.do_die || pm.do_recover)
        {
            KTimer kt(s, "pandemic_interact", 6.0 * t.n);
            launch_pandemic_interact(a, s.stream);
        }
        INC_CUDA(cudaGetLastError());
        if (pm.do_die) t.all_alive = false;
    }
    return INCERTO_OK;
}

// INCERTO_SPAWN_PANDEMIC_PEOPLE: examples/pandemic.rs:84-107 with the Philox contract
int32_t pandemic_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset)
{
    incerto_spawn_pandemic_params p;
    std::memcpy(&p, sp.params.data(), sizeof(p));
    if (row_offset != 0)
    {
        set_error("the pandemic device spawner must be the only spawner of its archetype");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    PandemicArgs a;
    std::memset(&a, 0, sizeof(a));
    a.pos = static_cast<uint32_t*>(t.col(p.position)->d);
    a.flags = t.d_flags;
    a.n = p.initial_population;
    a.grid_size = p.grid_size;
    a.bit_person = 1u << s.comps[p.person].marker_bit;
    a.bit_infected = 1u << s.comps[p.infected].marker_bit;
    a.thr_start_infected = bernoulli_threshold(p.chance_start_infected);
    a.rk_spawn = make_round_keys(make_key(s.seed, s.replica, TAG_PANDEMIC_SPAWN));
    KTimer kt(s, "pandemic_spawn", 5.0 * a.n);
    launch_pandemic_spawn(a, s.stream);
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

} // namespace incerto
