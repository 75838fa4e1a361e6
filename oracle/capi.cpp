// ORACLE (test infrastructure, NOT product code).
// Flat C interface over the CPU restatement so that tests/ (ctypes) and bench.py's cpu_baseline leg can drive it.
// Nothing under power-grid-model_b200/ links or loads this library.
#include "model.hpp"

#include <cstring>
#include <map>
#include <memory>

using namespace pgm_oracle;

namespace {

// results of variable size are handed back as named arrays
struct Bag {
    std::map<std::string, std::vector<int64_t>> i;
    std::map<std::string, std::vector<double>> d;
    std::string err;
};

enum Status : int { ok = 0, diverged = 1, singular = 2, other_error = 3 };

template <class F> int guarded(Bag* bag, F&& f) {
    try {
        f();
        return ok;
    } catch (IterationDiverge const& e) {
        if (bag != nullptr) bag->err = e.what();
        return diverged;
    } catch (SparseMatrixError const& e) {
        if (bag != nullptr) bag->err = e.what();
        return singular;
    } catch (std::exception const& e) {
        if (bag != nullptr) bag->err = e.what();
        return other_error;
    }
}

MathTopology make_topo(int64_t n_bus, double const* phase_shift, int64_t n_branch, int64_t const* branch_bus_idx,
                       int64_t n_fill, int64_t const* fill_in, int64_t const* sources_per_bus,
                       int64_t const* shunts_per_bus, int64_t const* load_gens_per_bus, int8_t const* load_gen_type) {
    MathTopology t;
    t.phase_shift.assign(phase_shift, phase_shift + n_bus);
    for (int64_t b = 0; b != n_branch; ++b) t.branch_bus_idx.push_back({branch_bus_idx[2 * b], branch_bus_idx[2 * b + 1]});
    for (int64_t f = 0; f != n_fill; ++f) t.fill_in.push_back({fill_in[2 * f], fill_in[2 * f + 1]});
    t.sources_per_bus.assign(sources_per_bus, sources_per_bus + n_bus + 1);
    t.shunts_per_bus.assign(shunts_per_bus, shunts_per_bus + n_bus + 1);
    t.load_gens_per_bus.assign(load_gens_per_bus, load_gens_per_bus + n_bus + 1);
    for (int64_t i = 0; i != t.n_load_gen(); ++i) t.load_gen_type.push_back(static_cast<LoadGenType>(load_gen_type[i]));
    return t;
}

template <int B> CMat<B> read_cmat(double const* p) { // row-major [r][c] of (re, im)
    CMat<B> m;
    for (int r = 0; r < B; ++r)
        for (int c = 0; c < B; ++c) m.m[r][c] = cplx{p[2 * (r * B + c)], p[2 * (r * B + c) + 1]};
    return m;
}
template <int B> void push_cvec(std::vector<double>& out, CVec<B> const& v) {
    for (int p = 0; p < B; ++p) {
        out.push_back(v.v[p].real());
        out.push_back(v.v[p].imag());
    }
}

template <int B>
void store_solver_output(Bag& bag, SolverOutput<B> const& so, std::string const& prefix = "") {
    auto& u = bag.d[prefix + "u"];
    auto& inj = bag.d[prefix + "bus_injection"];
    auto& br = bag.d[prefix + "branch"];
    auto& src = bag.d[prefix + "source"];
    auto& sh = bag.d[prefix + "shunt"];
    auto& lg = bag.d[prefix + "load_gen"];
    for (auto const& x : so.u) push_cvec<B>(u, x);
    for (auto const& x : so.bus_injection) push_cvec<B>(inj, x);
    for (auto const& x : so.branch) {
        push_cvec<B>(br, x.s_f);
        push_cvec<B>(br, x.s_t);
        push_cvec<B>(br, x.i_f);
        push_cvec<B>(br, x.i_t);
    }
    for (auto const& x : so.source) {
        push_cvec<B>(src, x.s);
        push_cvec<B>(src, x.i);
    }
    for (auto const& x : so.shunt) {
        push_cvec<B>(sh, x.s);
        push_cvec<B>(sh, x.i);
    }
    for (auto const& x : so.load_gen) {
        push_cvec<B>(lg, x.s);
        push_cvec<B>(lg, x.i);
    }
    bag.i[prefix + "num_iter"] = {so.num_iter};
}

template <int B>
void math_pf(Bag& bag, MathTopology const& topo, double const* branch_param, double const* shunt_param,
             double const* source_param, double const* source_u_ref, double const* s_injection, int method,
             double err_tol, int64_t max_iter) {
    MathParam<B> param;
    constexpr int BB2 = B * B * 2;
    for (Idx b = 0; b != topo.n_branch(); ++b) {
        BranchCalcParam<B> bp;
        for (int k = 0; k < 4; ++k) bp.value[k] = read_cmat<B>(branch_param + (b * 4 + k) * BB2);
        param.branch_param.push_back(bp);
    }
    for (Idx s = 0; s != topo.n_shunt(); ++s) param.shunt_param.push_back(read_cmat<B>(shunt_param + s * BB2));
    for (Idx s = 0; s != topo.n_source(); ++s) {
        param.source_param.push_back({cplx{source_param[4 * s], source_param[4 * s + 1]},
                                      cplx{source_param[4 * s + 2], source_param[4 * s + 3]}});
    }
    PowerFlowInput<B> input;
    for (Idx s = 0; s != topo.n_source(); ++s) input.source.push_back(cplx{source_u_ref[2 * s], source_u_ref[2 * s + 1]});
    for (Idx l = 0; l != topo.n_load_gen(); ++l) {
        CVec<B> v;
        for (int p = 0; p < B; ++p) v.v[p] = cplx{s_injection[(l * B + p) * 2], s_injection[(l * B + p) * 2 + 1]};
        input.s_injection.push_back(v);
    }
    auto topo_ptr = std::make_shared<MathTopology const>(topo);
    YBus<B> y_bus{topo_ptr, param};
    // expose the structure and admittances as well
    auto const& s = y_bus.structure();
    bag.i["row_indptr"] = s.row_indptr;
    bag.i["col_indices"] = s.col_indices;
    bag.i["row_indptr_lu"] = s.row_indptr_lu;
    bag.i["col_indices_lu"] = s.col_indices_lu;
    bag.i["diag_lu"] = s.diag_lu;
    bag.i["map_lu_y_bus"] = s.map_lu_y_bus;
    auto& adm = bag.d["admittance"];
    for (auto const& y : y_bus.admittance())
        for (int r = 0; r < B; ++r)
            for (int c = 0; c < B; ++c) {
                adm.push_back(y.m[r][c].real());
                adm.push_back(y.m[r][c].imag());
            }
    MathSolver<B> solver{topo};
    SolverOutput<B> const so =
        solver.run_power_flow(input, err_tol, max_iter, static_cast<CalculationMethod>(method), y_bus);
    store_solver_output<B>(bag, so);
}

struct ModelHandle {
    std::unique_ptr<Model> model;
};

} // namespace

extern "C" {

void* orc_bag_new() { return new Bag{}; }
void orc_bag_free(void* bag) { delete static_cast<Bag*>(bag); }
int orc_bag_get_i64(void* bag, char const* name, int64_t const** ptr, int64_t* n) {
    auto& m = static_cast<Bag*>(bag)->i;
    auto it = m.find(name);
    if (it == m.end()) return 1;
    *ptr = it->second.data();
    *n = static_cast<int64_t>(it->second.size());
    return 0;
}
int orc_bag_get_f64(void* bag, char const* name, double const** ptr, int64_t* n) {
    auto& m = static_cast<Bag*>(bag)->d;
    auto it = m.find(name);
    if (it == m.end()) return 1;
    *ptr = it->second.data();
    *n = static_cast<int64_t>(it->second.size());
    return 0;
}
char const* orc_bag_error(void* bag) { return static_cast<Bag*>(bag)->err.c_str(); }

// ---- sparse LU: block size N in {1,2,3,6}, real or complex; data column-major blocks (complex interleaved) ----
int orc_sparse_lu_solve(void* bag, int N, int is_complex, int64_t n, int64_t const* indptr, int64_t const* indices,
                        int64_t const* diag, double* data, double const* rhs, double* x, int use_pivot_perturbation,
                        int prefactorize_separately) {
    IdxVector const ip(indptr, indptr + n + 1);
    IdxVector const ix(indices, indices + ip.back());
    IdxVector const dg(diag, diag + n);
    auto run = [&]<class S, int NB>() {
        int64_t const nnz = ip.back();
        std::vector<S> d(nnz * NB * NB), r(n * NB), xx(n * NB);
        std::memcpy(d.data(), data, d.size() * sizeof(S));
        std::memcpy(r.data(), rhs, r.size() * sizeof(S));
        SparseLU<S, NB> solver{ip, ix, dg};
        typename SparseLU<S, NB>::PermArray perm;
        if (prefactorize_separately != 0) {
            solver.prefactorize(d, perm, use_pivot_perturbation != 0);
            solver.solve_with_prefactorized_matrix(d, perm, r, xx);
        } else {
            solver.prefactorize_and_solve(d, perm, r, xx, use_pivot_perturbation != 0);
        }
        std::memcpy(data, d.data(), d.size() * sizeof(S));
        std::memcpy(x, xx.data(), xx.size() * sizeof(S));
        if (bag != nullptr) {
            auto& pv = static_cast<Bag*>(bag)->i["perm"];
            pv.clear();
            for (auto const& bp : perm)
                for (int k = 0; k < NB; ++k) {
                    pv.push_back(bp.p[k]);
                }
            for (auto const& bp : perm)
                for (int k = 0; k < NB; ++k) {
                    pv.push_back(bp.q[k]);
                }
        }
    };
    return guarded(static_cast<Bag*>(bag), [&] {
        if (is_complex == 0) {
            switch (N) {
            case 1: run.template operator()<double, 1>(); break;
            case 2: run.template operator()<double, 2>(); break;
            case 3: run.template operator()<double, 3>(); break;
            case 6: run.template operator()<double, 6>(); break;
            default: throw PgmError{"unsupported block size"};
            }
        } else {
            switch (N) {
            case 1: run.template operator()<cplx, 1>(); break;
            case 3: run.template operator()<cplx, 3>(); break;
            default: throw PgmError{"unsupported block size"};
            }
        }
    });
}

int orc_min_degree(void* bag, int64_t n_vertex, int64_t const* keys, int64_t const* indptr, int64_t const* adj) {
    return guarded(static_cast<Bag*>(bag), [&] {
        ordering::Graph g;
        for (int64_t v = 0; v != n_vertex; ++v) g[keys[v]] = IdxVector(adj + indptr[v], adj + indptr[v + 1]);
        auto [alpha, fills] = ordering::minimum_degree_ordering(std::move(g));
        auto& b = *static_cast<Bag*>(bag);
        b.i["alpha"] = alpha;
        auto& f = b.i["fills"];
        for (auto [x, y] : fills) {
            f.push_back(x);
            f.push_back(y);
        }
    });
}

int orc_ybus_structure(void* bag, int64_t n_bus, int64_t n_branch, int64_t const* branch_bus_idx, int64_t n_fill,
                       int64_t const* fill_in, int64_t const* shunts_per_bus) {
    return guarded(static_cast<Bag*>(bag), [&] {
        std::vector<double> ps(n_bus, 0.0);
        IdxVector zeros(n_bus + 1, 0);
        MathTopology const topo = make_topo(n_bus, ps.data(), n_branch, branch_bus_idx, n_fill, fill_in, zeros.data(),
                                            shunts_per_bus, zeros.data(), nullptr);
        YBusStructure const s{topo};
        auto& b = *static_cast<Bag*>(bag);
        b.i["row_indptr"] = s.row_indptr;
        b.i["col_indices"] = s.col_indices;
        b.i["bus_entry"] = s.bus_entry;
        b.i["y_bus_entry_indptr"] = s.y_bus_entry_indptr;
        b.i["row_indptr_lu"] = s.row_indptr_lu;
        b.i["col_indices_lu"] = s.col_indices_lu;
        b.i["diag_lu"] = s.diag_lu;
        b.i["map_lu_y_bus"] = s.map_lu_y_bus;
        b.i["lu_transpose_entry"] = s.lu_transpose_entry;
        auto& et = b.i["y_bus_element_type"];
        auto& ei = b.i["y_bus_element_idx"];
        for (auto const& e : s.y_bus_element) {
            et.push_back(static_cast<int64_t>(e.element_type));
            ei.push_back(e.idx);
        }
    });
}

static void store_topology(Bag& b, std::vector<MathTopology> const& math, ComponentToMathCoupling const& coup) {
    b.i["n_math"] = {static_cast<int64_t>(math.size())};
    for (size_t g = 0; g != math.size(); ++g) {
        std::string const p = "g" + std::to_string(g) + ".";
        auto const& m = math[g];
        b.i[p + "slack_bus"] = {m.slack_bus};
        b.i[p + "is_radial"] = {m.is_radial ? 1 : 0};
        b.d[p + "phase_shift"] = m.phase_shift;
        auto& bb = b.i[p + "branch_bus_idx"];
        for (auto const& x : m.branch_bus_idx) {
            bb.push_back(x[0]);
            bb.push_back(x[1]);
        }
        auto& fi = b.i[p + "fill_in"];
        for (auto const& x : m.fill_in) {
            fi.push_back(x[0]);
            fi.push_back(x[1]);
        }
        b.i[p + "sources_per_bus"] = m.sources_per_bus;
        b.i[p + "shunts_per_bus"] = m.shunts_per_bus;
        b.i[p + "load_gens_per_bus"] = m.load_gens_per_bus;
        auto& lt = b.i[p + "load_gen_type"];
        for (auto t : m.load_gen_type) lt.push_back(static_cast<int64_t>(t));
        b.i[p + "voltage_regulators_per_load_gen"] = m.voltage_regulators_per_load_gen;
    }
    auto put = [&b](std::string const& name, std::vector<Idx2D> const& v) {
        auto& o = b.i[name];
        for (auto const& x : v) {
            o.push_back(x.group);
            o.push_back(x.pos);
        }
    };
    put("coup.node", coup.node);
    put("coup.branch", coup.branch);
    put("coup.shunt", coup.shunt);
    put("coup.load_gen", coup.load_gen);
    put("coup.source", coup.source);
    put("coup.voltage_regulator", coup.voltage_regulator);
    auto& b3 = b.i["coup.branch3"];
    for (auto const& [g, pos] : coup.branch3) {
        b3.push_back(g);
        b3.push_back(pos[0]);
        b3.push_back(pos[1]);
        b3.push_back(pos[2]);
    }
}

int orc_topology(void* bag, int64_t n_node, int64_t n_branch, int64_t const* branch_node_idx,
                 int8_t const* branch_connected, double const* branch_phase_shift, int64_t n_branch3,
                 int64_t const* branch3_node_idx, int8_t const* branch3_connected, double const* branch3_phase_shift,
                 int64_t n_source, int64_t const* source_node_idx, int8_t const* source_connected, int64_t n_shunt,
                 int64_t const* shunt_node_idx, int64_t n_load_gen, int64_t const* load_gen_node_idx,
                 int8_t const* load_gen_type) {
    return guarded(static_cast<Bag*>(bag), [&] {
        ComponentTopology ct;
        ComponentConnections cc;
        ct.n_node = n_node;
        for (int64_t b = 0; b != n_branch; ++b) {
            ct.branch_node_idx.push_back({branch_node_idx[2 * b], branch_node_idx[2 * b + 1]});
            cc.branch_connected.push_back({branch_connected[2 * b], branch_connected[2 * b + 1]});
            cc.branch_phase_shift.push_back(branch_phase_shift[b]);
        }
        for (int64_t b = 0; b != n_branch3; ++b) {
            ct.branch3_node_idx.push_back({branch3_node_idx[3 * b], branch3_node_idx[3 * b + 1], branch3_node_idx[3 * b + 2]});
            cc.branch3_connected.push_back({branch3_connected[3 * b], branch3_connected[3 * b + 1], branch3_connected[3 * b + 2]});
            cc.branch3_phase_shift.push_back({branch3_phase_shift[3 * b], branch3_phase_shift[3 * b + 1], branch3_phase_shift[3 * b + 2]});
        }
        ct.source_node_idx.assign(source_node_idx, source_node_idx + n_source);
        cc.source_connected.assign(source_connected, source_connected + n_source);
        ct.shunt_node_idx.assign(shunt_node_idx, shunt_node_idx + n_shunt);
        ct.load_gen_node_idx.assign(load_gen_node_idx, load_gen_node_idx + n_load_gen);
        for (int64_t i = 0; i != n_load_gen; ++i) ct.load_gen_type.push_back(static_cast<LoadGenType>(load_gen_type[i]));
        Topology topo{ct, cc};
        auto [math, coup] = topo.build_topology();
        store_topology(*static_cast<Bag*>(bag), math, coup);
    });
}

// math-level power flow on one sub-grid (the seam of tests/cpp_unit_tests/math_solver/test_math_solver_pf.hpp)
int orc_math_pf(void* bag, int sym, int64_t n_bus, double const* phase_shift, int64_t n_branch,
                int64_t const* branch_bus_idx, int64_t n_fill, int64_t const* fill_in, int64_t const* sources_per_bus,
                int64_t const* shunts_per_bus, int64_t const* load_gens_per_bus, int8_t const* load_gen_type,
                double const* branch_param, double const* shunt_param, double const* source_param,
                double const* source_u_ref, double const* s_injection, int method, double err_tol, int64_t max_iter) {
    return guarded(static_cast<Bag*>(bag), [&] {
        MathTopology const topo = make_topo(n_bus, phase_shift, n_branch, branch_bus_idx, n_fill, fill_in,
                                            sources_per_bus, shunts_per_bus, load_gens_per_bus, load_gen_type);
        if (sym != 0) {
            math_pf<1>(*static_cast<Bag*>(bag), topo, branch_param, shunt_param, source_param, source_u_ref, s_injection,
                       method, err_tol, max_iter);
        } else {
            math_pf<3>(*static_cast<Bag*>(bag), topo, branch_param, shunt_param, source_param, source_u_ref, s_injection,
                       method, err_tol, max_iter);
        }
    });
}

// ---- component-level model ----
void* orc_model_create(void* bag, double system_frequency, ModelInput const* input) {
    ModelHandle* h = nullptr;
    guarded(static_cast<Bag*>(bag), [&] {
        auto m = std::make_unique<Model>(system_frequency, *input);
        h = new ModelHandle{std::move(m)};
    });
    return h;
}
void orc_model_destroy(void* model) { delete static_cast<ModelHandle*>(model); }

// export the math model(s) the oracle derives from the components (topology, structure, parameters, PF input)
int orc_model_export_math(void* bag, void* model, int sym) {
    return guarded(static_cast<Bag*>(bag), [&] {
        auto& m = *static_cast<ModelHandle*>(model)->model;
        auto& b = *static_cast<Bag*>(bag);
        m.ensure_topology();
        std::vector<MathTopology> math;
        for (auto const& t : m.math_topology()) math.push_back(*t);
        store_topology(b, math, m.coupling());
        auto dump = [&]<int B>() {
            auto& ys = m.template prepared_y_bus<B>();
            auto const in = m.template power_flow_input<B>();
            for (size_t g = 0; g != ys.size(); ++g) {
                std::string const p = "g" + std::to_string(g) + ".";
                auto const& s = ys[g].structure();
                b.i[p + "row_indptr"] = s.row_indptr;
                b.i[p + "col_indices"] = s.col_indices;
                b.i[p + "bus_entry"] = s.bus_entry;
                b.i[p + "row_indptr_lu"] = s.row_indptr_lu;
                b.i[p + "col_indices_lu"] = s.col_indices_lu;
                b.i[p + "diag_lu"] = s.diag_lu;
                b.i[p + "map_lu_y_bus"] = s.map_lu_y_bus;
                b.i[p + "lu_transpose_entry"] = s.lu_transpose_entry;
                auto put_mat = [](std::vector<double>& o, CMat<B> const& y) {
                    for (int r = 0; r < B; ++r)
                        for (int c = 0; c < B; ++c) {
                            o.push_back(y.m[r][c].real());
                            o.push_back(y.m[r][c].imag());
                        }
                };
                auto& adm = b.d[p + "admittance"];
                for (auto const& y : ys[g].admittance()) put_mat(adm, y);
                auto& bp = b.d[p + "branch_param"];
                for (auto const& x : ys[g].param().branch_param)
                    for (int k = 0; k < 4; ++k) put_mat(bp, x.value[k]);
                auto& sp = b.d[p + "shunt_param"];
                for (auto const& x : ys[g].param().shunt_param) put_mat(sp, x);
                auto& srp = b.d[p + "source_param"];
                for (auto const& x : ys[g].param().source_param) {
                    srp.push_back(x.y1.real());
                    srp.push_back(x.y1.imag());
                    srp.push_back(x.y0.real());
                    srp.push_back(x.y0.imag());
                }
                auto& ur = b.d[p + "source_u_ref"];
                for (auto const& x : in[g].source) {
                    ur.push_back(x.real());
                    ur.push_back(x.imag());
                }
                auto& si = b.d[p + "s_injection"];
                for (auto const& x : in[g].s_injection) push_cvec<B>(si, x);
            }
        };
        if (sym != 0) {
            dump.template operator()<1>();
        } else {
            dump.template operator()<3>();
        }
    });
}

// batch (or single when update == nullptr) power flow. out points to a BatchOutput<1> or BatchOutput<3>.
// status[s] in {0 ok, 1 diverged, 2 singular, 3 other}; returns number of failed scenarios (or -1 on setup error)
int64_t orc_model_calculate(void* bag, void* model, int sym, int method, double err_tol, int64_t max_iter,
                            int64_t threading, int reuse_ic_factorization, BatchUpdate const* update, void const* out,
                            int64_t* n_iter, int32_t* status, int tap_strategy) {
    auto& m = *static_cast<ModelHandle*>(model)->model;
    auto* b = static_cast<Bag*>(bag);
    CalcOptions opt;
    opt.method = static_cast<CalculationMethod>(method);
    opt.err_tol = err_tol;
    opt.max_iter = max_iter;
    opt.threading = threading;
    opt.reuse_ic_factorization = reuse_ic_factorization != 0;
    opt.tap_strategy = tap_strategy; // PGM_TapChangingStrategy; not 0: automatic tap changer around every scenario's power flows
    int64_t failed = 0;
    auto run = [&]<int B>() {
        auto const& o = *static_cast<BatchOutput<B> const*>(out);
        if (update == nullptr) {
            int const st = guarded(b, [&] { m.template calculate<B>(opt, o, 0); });
            if (n_iter != nullptr) n_iter[0] = m.last_num_iter();
            if (status != nullptr) status[0] = st;
            failed = st == ok ? 0 : 1;
        } else {
            std::vector<std::string> const msgs = m.template batch_calculate<B>(opt, *update, o, n_iter);
            std::string all;
            for (size_t s = 0; s != msgs.size(); ++s) {
                int st = ok;
                if (!msgs[s].empty()) {
                    ++failed;
                    st = msgs[s].rfind("Iteration failed", 0) == 0 ? diverged
                         : msgs[s].rfind("Sparse matrix error", 0) == 0 ? singular
                                                                         : other_error;
                    all += "Error in batch #" + std::to_string(s) + ": " + msgs[s] + "\n";
                }
                if (status != nullptr) status[s] = st;
            }
            if (b != nullptr) b->err = all;
        }
    };
    int const st = guarded(b, [&] {
        if (sym != 0) {
            run.template operator()<1>();
        } else {
            run.template operator()<3>();
        }
    });
    return st == ok ? failed : -1;
}

// ranking of the regulated transformers by the automatic tap changer (tap_optimizer.hpp rank_transformers): rows of
// (kind: 0 transformer / 1 three-winding transformer, index within the kind, rank group) into bag.i["tap_rank"]
int orc_model_tap_rank(void* bag, void* model) {
    return guarded(static_cast<Bag*>(bag), [&] {
        auto& m = *static_cast<ModelHandle*>(model)->model;
        auto& out = static_cast<Bag*>(bag)->i["tap_rank"];
        out.clear();
        auto const groups = m.rank_transformers();
        for (size_t g = 0; g != groups.size(); ++g) {
            for (auto const& idx : groups[g]) {
                out.push_back(idx.group);
                out.push_back(idx.pos);
                out.push_back(static_cast<int64_t>(g));
            }
        }
    });
}

int64_t orc_hardware_concurrency() { return static_cast<int64_t>(std::thread::hardware_concurrency()); }

} // extern "C"
