// Benchmark input generator: the reference's fictional MV/LV distribution grid and load-profile batch
// (tests/benchmark_cpp/fictional_grid_generator.hpp:147-668; options from tests/benchmark_cpp/benchmark.cpp:257-263).
// The random streams (std::mt19937_64 + libstdc++ distributions, call order) are reproduced exactly so that seed 0 gives
// the grids BASELINE.md names; including the quirk that the LV ring list is not cleared between LV grids (:509-521).
#include "capi_common.hpp"
#include "components.hpp"

#include <random>
#include <vector>

using namespace pgmb;

struct pgmb_fictional_grid {
    pgmb_grid_option option{};
    int64_t n_lv_grid{};
    double ratio_lv_grid{};
    int64_t n_parallel_hv_mv_transformer{};
    std::mt19937_64 gen;
    ID id_gen{};
    std::vector<NodeInput> node;
    std::vector<TransformerInput> transformer;
    std::vector<LineInput> line;
    std::vector<SourceInput> source;
    std::vector<SymLoadGenInput> sym_load;
    std::vector<AsymLoadGenInput> asym_load;
    std::vector<ShuntInput> shunt;
    std::vector<ID> mv_ring, lv_ring;

    static void scale_cable(LineInput& line, double ratio) {
        line.r1 *= ratio;
        line.x1 *= ratio;
        line.c1 *= ratio;
        line.r0 *= ratio;
        line.x0 *= ratio;
        line.c0 *= ratio;
    }
    static TransformerInput transformer_template() {
        TransformerInput t{};
        t.i0_zero_sequence = kNaN;
        t.p0_zero_sequence = kNaN;
        t.uk_min = t.uk_max = t.pk_min = t.pk_max = kNaN;
        t.r_grounding_from = t.x_grounding_from = t.r_grounding_to = t.x_grounding_to = kNaN;
        return t;
    }

    void generate(uint32_t seed) {
        gen = std::mt19937_64{seed};
        id_gen = 0;
        auto& o = option;
        int64_t const n_mv_feeder_in = o.n_mv_feeder;
        int64_t total_mv_connection = o.n_mv_feeder * o.n_node_per_mv_feeder + 2;
        int64_t const node_per_lv_grid = o.n_lv_feeder * o.n_connection_per_lv_feeder * 2 + 1;
        if (total_mv_connection > o.n_node_total_specified) {
            n_lv_grid = 0;
            o.n_mv_feeder = (o.n_node_total_specified - 2) / o.n_node_per_mv_feeder;
            total_mv_connection = o.n_mv_feeder * o.n_node_per_mv_feeder;
        } else {
            n_lv_grid = (o.n_node_total_specified - total_mv_connection) / node_per_lv_grid;
        }
        if (n_lv_grid > total_mv_connection) o.n_mv_feeder = n_lv_grid / o.n_node_per_mv_feeder + 1;
        total_mv_connection = o.n_mv_feeder * o.n_node_per_mv_feeder;
        ratio_lv_grid = total_mv_connection > 0 ? static_cast<double>(n_lv_grid) / static_cast<double>(total_mv_connection) : 1.0;
        n_parallel_hv_mv_transformer = static_cast<int64_t>(static_cast<double>(n_mv_feeder_in) * 10.0 * 1.1 / 60.0) + 1;
        generate_mv_grid();
    }

    void generate_mv_grid() {
        auto const& o = option;
        ID const id_source_node = id_gen++;
        node.push_back({id_source_node, 150.0e3});
        source.push_back({id_gen++, id_source_node, 1, 1.05, kNaN, 2000e6, kNaN, kNaN});
        ID const id_mv_busbar = id_gen++;
        node.push_back({id_mv_busbar, 10.5e3});
        for (int64_t i = 0; i != n_parallel_hv_mv_transformer; ++i) {
            TransformerInput t = transformer_template();
            t.id = id_gen++;
            t.from_node = id_source_node;
            t.to_node = id_mv_busbar;
            t.from_status = 1;
            t.to_status = 1;
            t.u1 = 150.0e3;
            t.u2 = 10.5e3;
            t.sn = 60.0e6;
            t.uk = 0.203;
            t.pk = 200e3;
            t.i0 = 0.01;
            t.p0 = 40e3;
            t.winding_from = 1; // wye_n
            t.winding_to = 2;   // delta
            t.clock = 5;
            t.tap_side = 0;
            t.tap_pos = 0;
            t.tap_min = -10;
            t.tap_max = 10;
            t.tap_nom = 0;
            t.tap_size = 2.5e3;
            transformer.push_back(t);
            shunt.push_back({id_gen++, id_mv_busbar, 1, 0.0, 0.0, 0.0, -1.0 / 7.0});
        }
        SymLoadGenInput const mv_sym_load{0, 0, 1, 2 /*const_i*/, 0.8e6, 0.6e6};
        LineInput const mv_line{0, 0, 0, 1, 1, 0.063, 0.103, 0.4e-6, 0.0004, 0.275, 0.101, 0.66e-6, 0.0, 1e3};
        std::uniform_int_distribution<int64_t> load_type_gen{0, 2};
        std::uniform_real_distribution<double> scaling_gen{0.8 * 10.0 / static_cast<double>(o.n_node_per_mv_feeder),
                                                           1.2 * 10.0 / static_cast<double>(o.n_node_per_mv_feeder)};
        std::bernoulli_distribution lv_gen{ratio_lv_grid};
        for (int64_t i = 0; i < o.n_mv_feeder; i++) {
            ID prev_node_id = id_mv_busbar;
            for (int64_t j = 0; j < o.n_node_per_mv_feeder; ++j) {
                ID const current_node_id = id_gen++;
                node.push_back({current_node_id, 10.5e3});
                LineInput l = mv_line;
                l.id = id_gen++;
                l.from_node = prev_node_id;
                l.to_node = current_node_id;
                scale_cable(l, scaling_gen(gen));
                line.push_back(l);
                if (lv_gen(gen)) {
                    generate_lv_grid(current_node_id, 10.0 / static_cast<double>(o.n_node_per_mv_feeder));
                } else {
                    SymLoadGenInput s = mv_sym_load;
                    s.id = id_gen++;
                    s.node = current_node_id;
                    s.type = static_cast<IntS>(load_type_gen(gen));
                    double const sym_scale = scaling_gen(gen);
                    s.p_specified *= sym_scale;
                    s.q_specified *= sym_scale;
                    sym_load.push_back(s);
                }
                if (j == o.n_node_per_mv_feeder - 1) mv_ring.push_back(current_node_id);
                prev_node_id = current_node_id;
            }
        }
        if (mv_ring.size() > 1 && o.has_mv_ring != 0) {
            mv_ring.push_back(mv_ring.front());
            for (size_t k = 0; k + 1 < mv_ring.size(); ++k) {
                LineInput l = mv_line;
                l.id = id_gen++;
                l.from_node = mv_ring[k];
                l.to_node = mv_ring[k + 1];
                scale_cable(l, scaling_gen(gen));
                line.push_back(l);
            }
        }
    }

    void generate_lv_grid(ID mv_node, double mv_base_load) {
        auto const& o = option;
        ID const id_lv_busbar = id_gen++;
        node.push_back({id_lv_busbar, 400.0});
        TransformerInput t = transformer_template();
        t.id = id_gen++;
        t.from_node = mv_node;
        t.to_node = id_lv_busbar;
        t.from_status = 1;
        t.to_status = 1;
        t.u1 = 10.5e3;
        t.u2 = 420.0;
        t.sn = std::max(1500e3, mv_base_load * 1.2);
        t.uk = 0.06;
        t.pk = 8.8e3;
        t.i0 = 0.01;
        t.p0 = 1e3;
        t.winding_from = 2; // delta
        t.winding_to = 1;   // wye_n
        t.clock = 11;
        t.tap_side = 0;
        t.tap_pos = 3;
        t.tap_min = 5;
        t.tap_max = 1;
        t.tap_nom = 3;
        t.tap_size = 250.0;
        transformer.push_back(t);

        AsymLoadGenInput const lv_asym_load{0, 0, 1, 2 /*const_i*/, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        LineInput const lv_main_line{0, 0, 0, 1, 1, 0.206, 0.079, 0.72e-6, 0.0004, 0.94, 0.387, 0.36e-6, 0.0, 300.0};
        LineInput const lv_connection_line{0, 0, 0, 1, 1, 1.15, 0.096, 0.43e-6, 0.0004, 4.6, 0.408, 0.258e-6, 0.0, 80.0};
        std::uniform_int_distribution<int64_t> load_type_gen{0, 2};
        std::uniform_int_distribution<int64_t> load_phase_gen{0, 2};
        double const base_load = mv_base_load / static_cast<double>(o.n_lv_feeder * o.n_connection_per_lv_feeder) / 1.2;
        std::uniform_real_distribution<double> load_scaling_gen{0.8 * base_load, 1.2 * base_load};
        std::uniform_real_distribution<double> main_cable_gen{0.8 * 0.2 / static_cast<double>(o.n_connection_per_lv_feeder),
                                                              1.2 * 0.2 / static_cast<double>(o.n_connection_per_lv_feeder)};
        std::uniform_real_distribution<double> connection_cable_gen{5e-3, 20e-3};
        for (int64_t i = 0; i < o.n_lv_feeder; ++i) {
            ID prev_main_node_id = id_lv_busbar;
            for (int64_t j = 0; j < o.n_connection_per_lv_feeder; ++j) {
                ID const current_main_node_id = id_gen++;
                node.push_back({current_main_node_id, 400.0});
                ID const connection_node_id = id_gen++;
                node.push_back({connection_node_id, 400.0});
                LineInput main_line = lv_main_line;
                main_line.id = id_gen++;
                main_line.from_node = prev_main_node_id;
                main_line.to_node = current_main_node_id;
                scale_cable(main_line, main_cable_gen(gen));
                line.push_back(main_line);
                LineInput connection_line = lv_connection_line;
                connection_line.id = id_gen++;
                connection_line.from_node = current_main_node_id;
                connection_line.to_node = connection_node_id;
                scale_cable(connection_line, connection_cable_gen(gen));
                line.push_back(connection_line);
                AsymLoadGenInput a = lv_asym_load;
                a.id = id_gen++;
                a.node = connection_node_id;
                a.type = static_cast<IntS>(load_type_gen(gen));
                int64_t const phase = load_phase_gen(gen);
                double const apparent_power = load_scaling_gen(gen);
                a.p_specified[phase] = apparent_power * 0.8;
                a.q_specified[phase] = apparent_power * 0.6;
                asym_load.push_back(a);
                if (j == o.n_connection_per_lv_feeder - 1) lv_ring.push_back(current_main_node_id);
                prev_main_node_id = current_main_node_id;
            }
        }
        if (lv_ring.size() > 1 && o.has_lv_ring != 0) {
            lv_ring.push_back(lv_ring.front());
            for (size_t k = 0; k + 1 < lv_ring.size(); ++k) {
                LineInput l = lv_main_line;
                l.id = id_gen++;
                l.from_node = lv_ring[k];
                l.to_node = lv_ring[k + 1];
                scale_cable(l, main_cable_gen(gen));
                line.push_back(l);
            }
        }
    }

    // generate_batch_input / generate_load_series (:208-218, 644-668)
    void batch(int64_t batch_size, uint32_t seed, SymLoadGenUpdate* sym, AsymLoadGenUpdate* asym) {
        gen = std::mt19937_64{seed};
        std::uniform_real_distribution<double> scale{0.0, 1.0};
        int64_t const n_sym = static_cast<int64_t>(sym_load.size());
        for (int64_t b = 0; b != batch_size; ++b) {
            for (int64_t k = 0; k != n_sym; ++k) {
                SymLoadGenUpdate& u = sym[b * n_sym + k];
                u.id = sym_load[k].id;
                u.status = kNaIntS;
                u.p_specified = sym_load[k].p_specified * scale(gen);
                u.q_specified = sym_load[k].q_specified * scale(gen);
            }
        }
        std::uniform_real_distribution<double> scale2{0.0, 1.0};
        int64_t const n_asym = static_cast<int64_t>(asym_load.size());
        for (int64_t b = 0; b != batch_size; ++b) {
            for (int64_t k = 0; k != n_asym; ++k) {
                AsymLoadGenUpdate& u = asym[b * n_asym + k];
                u.id = asym_load[k].id;
                u.status = kNaIntS;
                for (int p = 0; p != 3; ++p) u.p_specified[p] = asym_load[k].p_specified[p] * scale2(gen);
                for (int p = 0; p != 3; ++p) u.q_specified[p] = asym_load[k].q_specified[p] * scale2(gen);
            }
        }
    }
};

extern "C" {

int pgmb_fictional_grid_create(const pgmb_grid_option* option, uint32_t seed, pgmb_fictional_grid** out) {
    return guarded([&] {
        if (option == nullptr || out == nullptr) throw InvalidArgument("null argument");
        if (option->n_node_per_mv_feeder <= 0 || option->n_lv_feeder <= 0 || option->n_connection_per_lv_feeder <= 0)
            throw InvalidArgument("grid option sizes must be positive");
        auto g = std::make_unique<pgmb_fictional_grid>();
        g->option = *option;
        g->generate(seed);
        *out = g.release();
    });
}
void pgmb_fictional_grid_destroy(pgmb_fictional_grid* grid) { delete grid; }

int pgmb_fictional_grid_get(pgmb_fictional_grid* grid, const char* component, const void** data, int64_t* n) {
    return guarded([&] {
        if (grid == nullptr || component == nullptr) throw InvalidArgument("null argument");
        std::string const c{component};
        auto give = [&](auto const& v) {
            *data = v.data();
            *n = static_cast<int64_t>(v.size());
        };
        if (c == "node") give(grid->node);
        else if (c == "line") give(grid->line);
        else if (c == "transformer") give(grid->transformer);
        else if (c == "shunt") give(grid->shunt);
        else if (c == "source") give(grid->source);
        else if (c == "sym_load") give(grid->sym_load);
        else if (c == "asym_load") give(grid->asym_load);
        else throw InvalidArgument("unknown component: " + c);
    });
}

int pgmb_fictional_grid_batch(pgmb_fictional_grid* grid, int64_t batch_size, uint32_t seed, void* sym_load_update,
                              void* asym_load_update) {
    return guarded([&] {
        if (grid == nullptr || batch_size < 0) throw InvalidArgument("invalid argument");
        if ((sym_load_update == nullptr && !grid->sym_load.empty()) || (asym_load_update == nullptr && !grid->asym_load.empty()))
            throw InvalidArgument("null update buffer");
        grid->batch(batch_size, seed, static_cast<SymLoadGenUpdate*>(sym_load_update),
                    static_cast<AsymLoadGenUpdate*>(asym_load_update));
    });
}

} // extern "C"
