// Native Python module `pypendulum` over the C++ facade (host/optimization.hpp -> include/cpmpc.h -> libcpmpc.so).
//
// It is the binding a maintainer of the reference would keep: the class / method / attribute list of
// wrapper/wrapper.cc:40-102 (nanobind there; pybind11 here because nanobind is an empty submodule of the reference
// and not installed), written against the SAME C++ names — pendulum::SingleCartPoleParams, OptimizationParams,
// SingleCartPoleState, OptimizationOutputs, Optimization, Vector2, Simulator — which host/optimization.hpp provides
// on top of the GPU library.  model/scratch.py:14-40 runs against the built module with only its sys.path line
// changed (tests/test_gpu_parity.py::test_native_pypendulum_runs_scratch_py).
//
// Built in-tree by cart_pole_mpc_b200/build.py into cart_pole_mpc_b200/native/pypendulum*.so (g++, no nvcc).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "optimization.hpp"

namespace py = pybind11;

namespace pendulum {

static void wrap_everything(py::module_& m) {
  py::class_<SingleCartPoleParams>(m, "SingleCartPoleParams")
      .def(py::init<>())
      .def(py::init<double, double, double, double, double, double, double, double, double>(), py::arg("m_b"),
           py::arg("m_1"), py::arg("l_1"), py::arg("g"), py::arg("mu_b"), py::arg("v_mu_b"), py::arg("c_d_1"),
           py::arg("x_s"), py::arg("k_s"))
      .def_readwrite("m_b", &SingleCartPoleParams::m_b)
      .def_readwrite("m_1", &SingleCartPoleParams::m_1)
      .def_readwrite("l_1", &SingleCartPoleParams::l_1)
      .def_readwrite("g", &SingleCartPoleParams::g)
      .def_readwrite("mu_b", &SingleCartPoleParams::mu_b)
      .def_readwrite("v_mu_b", &SingleCartPoleParams::v_mu_b)
      .def_readwrite("c_d_1", &SingleCartPoleParams::c_d_1)
      .def_readwrite("x_s", &SingleCartPoleParams::x_s)
      .def_readwrite("k_s", &SingleCartPoleParams::k_s);

  py::class_<OptimizationParams>(m, "OptimizationParams")
      .def(py::init<>())
      .def_readwrite("control_dt", &OptimizationParams::control_dt)
      .def_readwrite("window_length", &OptimizationParams::window_length)
      .def_readwrite("state_spacing", &OptimizationParams::state_spacing)
      .def_readwrite("max_iterations", &OptimizationParams::max_iterations)
      .def_readwrite("relative_exit_tol", &OptimizationParams::relative_exit_tol)
      .def_readwrite("absolute_first_derivative_tol", &OptimizationParams::absolute_first_derivative_tol)
      .def_readwrite("equality_penalty_initial", &OptimizationParams::equality_penalty_initial)
      .def_readwrite("u_guess_sinusoid_amplitude", &OptimizationParams::u_guess_sinusoid_amplitude)
      .def_readwrite("u_cost_weight", &OptimizationParams::u_cost_weight)
      .def_readwrite("u_derivative_cost_weight", &OptimizationParams::u_derivative_cost_weight)
      .def_readwrite("b_x_final_cost_weight", &OptimizationParams::b_x_final_cost_weight)
      .def_readwrite("th_final_cost_weight", &OptimizationParams::th_final_cost_weight)
      .def_readwrite("b_x_dot_final_cost_weight", &OptimizationParams::b_x_dot_final_cost_weight)
      .def_readwrite("th_dot_final_cost_weight", &OptimizationParams::th_dot_final_cost_weight);

  py::class_<SingleCartPoleState>(m, "SingleCartPoleState")
      .def(py::init<double, double, double, double>())
      .def_readwrite("b_x", &SingleCartPoleState::b_x)
      .def_readwrite("th_1", &SingleCartPoleState::th_1)
      .def_readwrite("b_x_dot", &SingleCartPoleState::b_x_dot)
      .def_readwrite("th_1_dot", &SingleCartPoleState::th_1_dot);

  py::class_<OptimizationOutputs>(m, "OptimizationOutputs")
      .def("solver_summary", [](const OptimizationOutputs& self) { return self.solver_outputs.ToString(); })
      .def_readonly("u", &OptimizationOutputs::u)
      .def_readonly("predicted_states", &OptimizationOutputs::predicted_states);

  py::class_<Optimization>(m, "Optimization")
      .def(py::init<const OptimizationParams&>())
      .def("step", &Optimization::Step)
      .def("set_previous_solution", &Optimization::SetPreviousSolution);

  py::class_<Vector2>(m, "Vector2").def(py::init([](double x, double y) { return Vector2{x, y}; }));

  py::class_<Simulator>(m, "Simulator")
      .def(py::init<>())
      .def("step", &Simulator::Step)
      .def("get_state", &Simulator::GetState);
}

}  // namespace pendulum

PYBIND11_MODULE(pypendulum, m) { pendulum::wrap_everything(m); }
