// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product path.
//
// CPU restatement of the reference's *sequential, un-fused, un-tiled* evaluation of a
// stencil sequence, rendered per program by oracle/reference_oracle.py:
//   * field layout          -- ref template_tiling.cpp:59-79  (i fastest, halo included, no padding)
//   * input initialisation  -- ref template_tiling.cpp:62-64,209-210 (srand(0); std::rand per cell,
//                              fields in declaration order) then make_periodic (:220-221)
//   * make_periodic         -- ref template_tiling.cpp:136-159 (i, then j, then k; sequential copies)
//   * evaluation order      -- ref template_tiling.cpp:162-177 (every stencil over the whole interior
//                              in sequence order, make_periodic after EACH stencil)
// The stencil bodies are the catalogue strings pasted verbatim, so the floating-point evaluation
// order is the reference's.  Build strictly (-O2 -ffp-contract=off, no -ffast-math) for the
// bit-exact comparison.  Pinned against the reference's own compute_reference (oracle/_ref, built by
// oracle/build_ref.py) and against the committed golden digests in tests/golden/.
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr int X = {{X}}, Y = {{Y}}, Z = {{Z}};
constexpr int HX = {{HX}}, HY = {{HY}}, HZ = {{HZ}};
constexpr long ROW = X + 2 * HX;
constexpr long PLANE = ROW * (Y + 2 * HY);
constexpr long CELLS = PLANE * (Z + 2 * HZ);

struct field {
    double* mem;
    double& operator()(int i, int j, int k) const { return mem[i + j * ROW + k * PLANE]; }
};
struct const_field {
    const double* mem;
    const double& operator()(int i, int j, int k) const { return mem[i + j * ROW + k * PLANE]; }
};

void make_periodic(field data) {
    for (int k = 0; k < Z + 2 * HZ; ++k)
        for (int j = 0; j < Y + 2 * HY; ++j)
            for (int i = 0; i < HX; ++i) {
                data(i, j, k) = data(i + X, j, k);
                data(i + X + HX, j, k) = data(i + HX, j, k);
            }
    for (int k = 0; k < Z + 2 * HZ; ++k)
        for (int j = 0; j < HY; ++j)
            for (int i = 0; i < X + 2 * HX; ++i) {
                data(i, j, k) = data(i, j + Y, k);
                data(i, j + Y + HY, k) = data(i, j + HY, k);
            }
    for (int k = 0; k < HZ; ++k)
        for (int j = 0; j < Y + 2 * HY; ++j)
            for (int i = 0; i < X + 2 * HX; ++i) {
                data(i, j, k) = data(i, j, k + Z);
                data(i, j, k + Z + HZ) = data(i, j, k + HZ);
            }
}

const char* const input_names[] = { {% for name in INPUTS %}"{{name}}", {% endfor %}nullptr };
const char* const output_names[] = { {% for name in OUTPUTS %}"{{name}}", {% endfor %}nullptr };

}  // namespace

extern "C" {

void oracle_info(int* dims, int* n_inputs, int* n_outputs) {
    dims[0] = X; dims[1] = Y; dims[2] = Z; dims[3] = HX; dims[4] = HY; dims[5] = HZ;
    *n_inputs = {{INPUTS|length}};
    *n_outputs = {{OUTPUTS|length}};
}

const char* oracle_field_name(int is_output, int idx) {
    return is_output ? output_names[idx] : input_names[idx];
}

void oracle_make_periodic(double* data) { make_periodic(field{data}); }

// srand(seed) once, then every cell of every input from rand(), fields in the given order
void oracle_fill_inputs(double* const* inputs, int count, unsigned seed) {
    std::srand(seed);
    for (int n = 0; n < count; ++n)
        for (long c = 0; c < CELLS; ++c) inputs[n][c] = std::rand();
    for (int n = 0; n < count; ++n) make_periodic(field{inputs[n]});
}

// inputs: periodic fields in INPUTS order; outputs: caller-allocated fields in OUTPUTS order
void oracle_run(const double* const* inputs, double* const* outputs) {
    {% for name in INPUTS %}
    const const_field {{name}}{inputs[{{loop.index0}}]};{% endfor %}
    {% for name in OUTPUTS %}
    const field {{name}}{outputs[{{loop.index0}}]};
    std::memset({{name}}.mem, 0, sizeof(double) * CELLS);{% endfor %}
    {% for name in TEMPS %}
    std::vector<double> store_{{name}}(CELLS, 0.0);
    const field {{name}}{store_{{name}}.data()};{% endfor %}
    {% for stencil in STENCILS %}
    for (int k = HZ; k < Z + HZ; ++k)
        for (int j = HY; j < Y + HY; ++j)
            for (int i = HX; i < X + HX; ++i) {
                {{stencil.LAMBDA}}
                {{stencil.NAME}}(i, j, k) = res;
            }
    make_periodic({{stencil.NAME}});
    {% endfor %}
}

}  // extern "C"
