// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the reference block-sparse LU solver:
//   power_grid_model/math_solver/sparse_lu_solver.hpp
//     perturb_pivot_if_needed :39-48, DenseLUFactor::factorize_block_in_place :86-165,
//     triangular_solve_inplace :171-200, SparseLUSolver::prefactorize :346-495,
//     solve_with_refinement :514-538, iterate_and_backward_error :580-622,
//     initialize_pivot_perturbation :624-649, find_entry :734-748, solve_once :769-827.
// Eigen semantics restated by hand: blocks are column-major, maxCoeff scans column-major and
// keeps the first maximum, PermutationMatrix{ind}: (P v)[ind[i]] = v[i], (M Q)[:, i] = M[:, ind[i]].
// Layout: a block matrix is a flat std::vector<S>, block k at [k*N*N, (k+1)*N*N), element (r, c) at c*N + r.
// N == 1 is the reference's "scalar" specialisation (no block permutation, no dense factorisation).
#pragma once

#include "tensor.hpp"

#include <optional>

namespace pgm_oracle {

constexpr double epsilon = std::numeric_limits<double>::epsilon();
constexpr double epsilon_perturbation = 1e-13;
constexpr double cap_back_error_denominator = 1e-4;

template <class S> inline void perturb_pivot_if_needed(double perturb_threshold, S& value, double& abs_value,
                                                       bool& has_pivot_perturbation) {
    if (abs_value < perturb_threshold) {
        S const scale = (abs_value == 0.0) ? S{1.0} : (value / abs_value);
        value = scale * perturb_threshold;
        has_pivot_perturbation = true;
        abs_value = perturb_threshold;
    }
}

template <int N> struct BlockPerm {
    int8_t p[N];
    int8_t q[N];
};

enum class Side { left, right };
enum class Factor { lower, upper };

template <class S, int N> struct DenseLU {
    static S& at(S* m, int r, int c) { return m[c * N + r]; }
    static S const& at(S const* m, int r, int c) { return m[c * N + r]; }

    // sparse_lu_solver.hpp:86-165
    static void factorize_block_in_place(S* matrix, BlockPerm<N>& block_perm, double perturb_threshold,
                                         bool use_pivot_perturbation, bool& has_pivot_perturbation) {
        int8_t row_transpositions[N]{};
        int8_t col_transpositions[N]{};
        double max_pivot{};

        for (int pivot = 0; pivot != N; ++pivot) {
            // cwiseAbs2().maxCoeff(&r,&c) over the bottom-right corner: column-major visit, first maximum wins
            int row_biggest = pivot;
            int col_biggest = pivot;
            double biggest_score = abs2(at(matrix, pivot, pivot));
            for (int c = pivot; c != N; ++c) {
                for (int r = pivot; r != N; ++r) {
                    double const score = abs2(at(matrix, r, c));
                    if (score > biggest_score) {
                        biggest_score = score;
                        row_biggest = r;
                        col_biggest = c;
                    }
                }
            }
            if (biggest_score == 0.0 && !use_pivot_perturbation) {
                for (int rest = pivot; rest != N; ++rest) {
                    row_transpositions[rest] = static_cast<int8_t>(rest);
                    col_transpositions[rest] = static_cast<int8_t>(rest);
                }
                break;
            }
            double abs_pivot = std::sqrt(biggest_score);
            perturb_pivot_if_needed(perturb_threshold, at(matrix, row_biggest, col_biggest), abs_pivot,
                                    has_pivot_perturbation);
            max_pivot = std::max(max_pivot, abs_pivot);

            row_transpositions[pivot] = static_cast<int8_t>(row_biggest);
            col_transpositions[pivot] = static_cast<int8_t>(col_biggest);
            if (pivot != row_biggest) {
                for (int c = 0; c != N; ++c) std::swap(at(matrix, pivot, c), at(matrix, row_biggest, c));
            }
            if (pivot != col_biggest) {
                for (int r = 0; r != N; ++r) std::swap(at(matrix, r, pivot), at(matrix, r, col_biggest));
            }
            if (pivot < N - 1) {
                for (int r = pivot + 1; r != N; ++r) at(matrix, r, pivot) /= at(matrix, pivot, pivot);
                for (int c = pivot + 1; c != N; ++c)
                    for (int r = pivot + 1; r != N; ++r)
                        at(matrix, r, c) -= at(matrix, r, pivot) * at(matrix, pivot, c);
            }
        }
        // accumulate permutations :147-155
        for (int i = 0; i != N; ++i) {
            block_perm.p[i] = static_cast<int8_t>(i);
            block_perm.q[i] = static_cast<int8_t>(i);
        }
        for (int pivot = N - 1; pivot != -1; --pivot) std::swap(block_perm.p[pivot], block_perm.p[row_transpositions[pivot]]);
        for (int pivot = 0; pivot != N; ++pivot) std::swap(block_perm.q[pivot], block_perm.q[col_transpositions[pivot]]);

        double const pivot_threshold = has_pivot_perturbation ? 0.0 : epsilon * max_pivot;
        for (int pivot = 0; pivot != N; ++pivot) {
            if (cabs(at(matrix, pivot, pivot)) < pivot_threshold || !is_normal(at(matrix, pivot, pivot))) {
                throw SparseMatrixError{};
            }
        }
    }

    // rhs is N x M column-major (M = N for a block, M = 1 for a vector). sparse_lu_solver.hpp:171-200
    template <Side side, Factor factor, int M> static void triangular_solve_inplace(S const* lu, S* rhs) {
        constexpr bool forward = (side == Side::left) == (factor == Factor::lower);
        auto R = [&](int r, int c) -> S& { return rhs[c * N + r]; };
        for (int step = 0; step != N; ++step) {
            int const index = forward ? step : N - 1 - step;
            for (int prev_step = 0; prev_step != step; ++prev_step) {
                int const prev = forward ? prev_step : N - 1 - prev_step;
                if constexpr (side == Side::left) {
                    for (int c = 0; c != M; ++c) R(index, c) -= at(lu, index, prev) * R(prev, c);
                } else {
                    for (int r = 0; r != N; ++r) R(r, index) -= at(lu, prev, index) * R(r, prev);
                }
            }
            if constexpr (factor == Factor::upper) {
                if constexpr (side == Side::left) {
                    for (int c = 0; c != M; ++c) R(index, c) /= at(lu, index, index);
                } else {
                    for (int r = 0; r != N; ++r) R(r, index) /= at(lu, index, index);
                }
            }
        }
    }
    // out = P * in (rows): out row p[i] = in row i ; M columns
    template <int M> static void perm_rows(int8_t const* p, S* blk) {
        S tmp[N * M];
        for (int i = 0; i != N * M; ++i) tmp[i] = blk[i];
        for (int c = 0; c != M; ++c)
            for (int i = 0; i != N; ++i) blk[c * N + p[i]] = tmp[c * N + i];
    }
    // out = in * Q (cols): out col i = in col q[i]
    static void perm_cols(int8_t const* q, S* blk) {
        S tmp[N * N];
        for (int i = 0; i != N * N; ++i) tmp[i] = blk[i];
        for (int i = 0; i != N; ++i)
            for (int r = 0; r != N; ++r) blk[i * N + r] = tmp[q[i] * N + r];
    }
};

template <class S, int N> class SparseLU {
  public:
    static constexpr bool is_block = N > 1;
    static constexpr int NN = N * N;
    static constexpr Idx max_iterative_refinement = 5;
    using LU = DenseLU<S, N>;
    using PermArray = std::vector<BlockPerm<N>>;

    SparseLU(IdxVector const& row_indptr, IdxVector const& col_indices, IdxVector const& diag_lu)
        : size_{static_cast<Idx>(row_indptr.size()) - 1},
          nnz_{row_indptr.back()},
          row_indptr_{&row_indptr},
          col_indices_{&col_indices},
          diag_lu_{&diag_lu} {}

    void prefactorize_and_solve(std::vector<S>& data, PermArray& perm, std::vector<S> const& rhs, std::vector<S>& x,
                                bool use_pivot_perturbation = false) {
        prefactorize(data, perm, use_pivot_perturbation);
        solve_with_prefactorized_matrix(data, perm, rhs, x);
    }
    void solve_with_prefactorized_matrix(std::vector<S> const& data, PermArray const& perm, std::vector<S> const& rhs,
                                         std::vector<S>& x) {
        if (has_pivot_perturbation_) {
            solve_with_refinement(data, perm, rhs, x);
        } else {
            solve_once(data, perm, rhs, x);
        }
    }

    // sparse_lu_solver.hpp:346-495
    void prefactorize(std::vector<S>& lu, PermArray& perm, bool use_pivot_perturbation = false) {
        auto const& indptr = *row_indptr_;
        auto const& indices = *col_indices_;
        auto const& diag = *diag_lu_;
        reset_matrix_cache();
        if (use_pivot_perturbation) {
            initialize_pivot_perturbation(lu);
        }
        double const perturb_threshold = epsilon_perturbation * matrix_norm_;
        if constexpr (is_block) {
            perm.resize(size_);
        }
        IdxVector col_position_idx(indptr.begin(), indptr.end() - 1);

        for (Idx pivot_row_col = 0; pivot_row_col != size_; ++pivot_row_col) {
            Idx const pivot_idx = diag[pivot_row_col];
            S* const pivot = &lu[pivot_idx * NN];
            if constexpr (is_block) {
                LU::factorize_block_in_place(pivot, perm[pivot_row_col], perturb_threshold, use_pivot_perturbation,
                                             has_pivot_perturbation_);
            } else {
                if (use_pivot_perturbation) {
                    double abs_pivot = cabs(*pivot);
                    perturb_pivot_if_needed(perturb_threshold, *pivot, abs_pivot, has_pivot_perturbation_);
                }
                if (!is_normal(*pivot)) {
                    throw SparseMatrixError{};
                }
            }
            if constexpr (is_block) {
                BlockPerm<N> const& bp = perm[pivot_row_col];
                // permute already-computed L (rows) left of the pivot and U (cols) above it :399-415
                for (Idx l_idx = indptr[pivot_row_col]; l_idx < pivot_idx; ++l_idx) {
                    LU::template perm_rows<N>(bp.p, &lu[l_idx * NN]);
                    Idx const u_row = indices[l_idx];
                    Idx const u_idx = col_position_idx[u_row];
                    LU::perm_cols(bp.q, &lu[u_idx * NN]);
                    ++col_position_idx[u_row];
                }
                // U blocks right of the pivot :420-429
                for (Idx u_idx = pivot_idx + 1; u_idx < indptr[pivot_row_col + 1]; ++u_idx) {
                    S* const u = &lu[u_idx * NN];
                    LU::template perm_rows<N>(bp.p, u);
                    LU::template triangular_solve_inplace<Side::left, Factor::lower, N>(pivot, u);
                }
            }
            // L blocks below the pivot + Schur complement :437-487
            for (Idx l_ref_idx = pivot_idx + 1; l_ref_idx < indptr[pivot_row_col + 1]; ++l_ref_idx) {
                Idx const l_row = indices[l_ref_idx];
                Idx const l_idx = col_position_idx[l_row];
                S* const l = &lu[l_idx * NN];
                if constexpr (is_block) {
                    LU::perm_cols(perm[pivot_row_col].q, l);
                    LU::template triangular_solve_inplace<Side::right, Factor::upper, N>(pivot, l);
                } else {
                    *l = *l / *pivot;
                }
                Idx a_idx = l_idx;
                for (Idx u_idx = pivot_idx + 1; u_idx < indptr[pivot_row_col + 1]; ++u_idx) {
                    Idx const u_col = indices[u_idx];
                    a_idx = find_entry(l_row, u_col, a_idx + 1, indptr[l_row + 1]);
                    // lu[a] -= dot(l, u)
                    S const* const u = &lu[u_idx * NN];
                    S* const av = &lu[a_idx * NN];
                    for (int c = 0; c != N; ++c)
                        for (int r = 0; r != N; ++r) {
                            S s = l[0 * N + r] * u[c * N + 0];
                            for (int k = 1; k != N; ++k) s += l[k * N + r] * u[c * N + k];
                            av[c * N + r] -= s;
                        }
                }
                ++col_position_idx[l_row];
            }
            ++col_position_idx[pivot_row_col];
        }
        if (!has_pivot_perturbation_) {
            reset_matrix_cache();
        }
    }

    bool has_pivot_perturbation() const { return has_pivot_perturbation_; }

    // sparse_lu_solver.hpp:769-827
    void solve_once(std::vector<S> const& lu, PermArray const& perm, std::vector<S> const& rhs,
                    std::vector<S>& x) const {
        auto const& indptr = *row_indptr_;
        auto const& indices = *col_indices_;
        auto const& diag = *diag_lu_;
        auto sub_dot = [](S* xr, S const* blk, S const* xc) { // xr -= dot(blk, xc)
            for (int r = 0; r != N; ++r) {
                S s = blk[0 * N + r] * xc[0];
                for (int k = 1; k != N; ++k) s += blk[k * N + r] * xc[k];
                xr[r] -= s;
            }
        };
        for (Idx row = 0; row != size_; ++row) {
            S tmp[N];
            for (int i = 0; i != N; ++i) tmp[i] = rhs[row * N + i]; // rhs may alias x
            if constexpr (is_block) {
                for (int i = 0; i != N; ++i) x[row * N + perm[row].p[i]] = tmp[i];
            } else {
                x[row] = tmp[0];
            }
            for (Idx l_idx = indptr[row]; l_idx < diag[row]; ++l_idx) {
                sub_dot(&x[row * N], &lu[l_idx * NN], &x[indices[l_idx] * N]);
            }
            if constexpr (is_block) {
                LU::template triangular_solve_inplace<Side::left, Factor::lower, 1>(&lu[diag[row] * NN], &x[row * N]);
            }
        }
        for (Idx row = size_ - 1; row != -1; --row) {
            for (Idx u_idx = indptr[row + 1] - 1; u_idx > diag[row]; --u_idx) {
                sub_dot(&x[row * N], &lu[u_idx * NN], &x[indices[u_idx] * N]);
            }
            if constexpr (is_block) {
                LU::template triangular_solve_inplace<Side::left, Factor::upper, 1>(&lu[diag[row] * NN], &x[row * N]);
            } else {
                x[row] = x[row] / lu[diag[row]];
            }
        }
        if constexpr (is_block) {
            for (Idx row = 0; row != size_; ++row) {
                S tmp[N];
                for (int i = 0; i != N; ++i) tmp[i] = x[row * N + i];
                for (int i = 0; i != N; ++i) x[row * N + perm[row].q[i]] = tmp[i];
            }
        }
    }

  private:
    static constexpr Idx linear_search_threshold = 16;
    Idx size_;
    Idx nnz_;
    IdxVector const* row_indptr_;
    IdxVector const* col_indices_;
    IdxVector const* diag_lu_;
    bool has_pivot_perturbation_{false};
    double matrix_norm_{};
    std::optional<std::vector<S>> original_matrix_;

    Idx find_entry(Idx /*row*/, Idx col, Idx begin_idx, Idx end_idx) const {
        auto const first = col_indices_->begin() + begin_idx;
        auto const last = col_indices_->begin() + end_idx;
        auto const found =
            end_idx - begin_idx < linear_search_threshold ? std::find(first, last, col) : std::lower_bound(first, last, col);
        if (found == last || *found != col) {
            throw PgmError{"sparse LU pattern is not closed under fill-in"};
        }
        return static_cast<Idx>(found - col_indices_->begin());
    }
    void reset_matrix_cache() {
        has_pivot_perturbation_ = false;
        matrix_norm_ = 0.0;
        original_matrix_.reset();
    }
    // :624-649
    void initialize_pivot_perturbation(std::vector<S> const& data) {
        auto const& indptr = *row_indptr_;
        auto const& indices = *col_indices_;
        original_matrix_ = data;
        matrix_norm_ = 0.0;
        for (Idx row = 0; row != size_; ++row) {
            double row_norm = 0.0;
            for (Idx idx = indptr[row]; idx != indptr[row + 1]; ++idx) {
                if (indices[idx] == row) continue;
                // cabs(block).rowwise().sum().maxCoeff()
                double block_norm = 0.0;
                for (int r = 0; r != N; ++r) {
                    double s = 0.0;
                    for (int c = 0; c != N; ++c) s += cabs(data[idx * NN + c * N + r]);
                    block_norm = (r == 0) ? s : std::max(block_norm, s);
                }
                row_norm += block_norm;
            }
            matrix_norm_ = std::max(matrix_norm_, row_norm);
        }
    }
    // :514-538, 546-622
    void solve_with_refinement(std::vector<S> const& data, PermArray const& perm, std::vector<S> const& rhs_in,
                               std::vector<S>& x) {
        auto const& indptr = *row_indptr_;
        auto const& indices = *col_indices_;
        auto const& original = original_matrix_.value();
        std::vector<S> const rhs = rhs_in;
        std::fill(x.begin(), x.end(), S{0.0});
        std::vector<S> residual = rhs;
        std::vector<S> dx(x.size(), S{0.0});
        double backward_error = std::numeric_limits<double>::max();
        Idx num_iter = 0;
        while (backward_error > epsilon_perturbation) {
            if (num_iter++ == max_iterative_refinement + 1) {
                throw SparseMatrixError{};
            }
            solve_once(data, perm, residual, dx);
            // iterate_and_backward_error
            std::vector<double> denom(size_ * N);
            double max_denominator = 0.0;
            for (Idx row = 0; row != size_; ++row) {
                for (int r = 0; r != N; ++r) denom[row * N + r] = cabs(rhs[row * N + r]);
                for (Idx idx = indptr[row]; idx != indptr[row + 1]; ++idx) {
                    Idx const col = indices[idx];
                    for (int r = 0; r != N; ++r) {
                        double s = cabs(original[idx * NN + 0 * N + r]) * cabs(x[col * N + 0]);
                        for (int k = 1; k != N; ++k) s += cabs(original[idx * NN + k * N + r]) * cabs(x[col * N + k]);
                        denom[row * N + r] += s;
                    }
                }
                for (int r = 0; r != N; ++r) max_denominator = std::max(max_denominator, denom[row * N + r]);
            }
            double const min_denominator = cap_back_error_denominator * max_denominator;
            double max_berr = 0.0;
            for (Idx i = 0; i != size_ * N; ++i) {
                double const d = std::max(denom[i], min_denominator);
                max_berr = std::max(max_berr, cabs(residual[i]) / d);
                x[i] += dx[i];
            }
            backward_error = max_berr;
            // calculate_residual
            for (Idx row = 0; row != size_; ++row) {
                for (int r = 0; r != N; ++r) residual[row * N + r] = rhs[row * N + r];
                for (Idx idx = indptr[row]; idx != indptr[row + 1]; ++idx) {
                    Idx const col = indices[idx];
                    for (int r = 0; r != N; ++r) {
                        S s = original[idx * NN + 0 * N + r] * x[col * N + 0];
                        for (int k = 1; k != N; ++k) s += original[idx * NN + k * N + r] * x[col * N + k];
                        residual[row * N + r] -= s;
                    }
                }
            }
        }
    }
};

} // namespace pgm_oracle
