#!/bin/bash
# Source ranges of the solve kernel for profiles/ncu_lines.py --ranges (line numbers of the tree this file is committed
# with): where the samples of an `ncu --set full --import-source on` capture fall, by solver phase and function.
#   bash profiles/ncu_ranges.sh gpurun_out/<capture>.ncu-rep
R="fastmath.cuh:100-133=rcp,fastmath.cuh:134-148=sqrt_rsqrt,fastmath.cuh:149-202=sincos,fastmath.cuh:203-260=tanh"
R="$R,dynamics.cuh:152-193=dyn_eval,dynamics.cuh:194-258=dyn_eval_jac,dynamics.cuh:259-291=rk4_step,dynamics.cuh:292-370=rk4_step_jac (chain rule)"
R="$R,dynamics.cuh:100-151=load constants+mod_pi"
R="$R,solver_kernel.cuh:139-169=control costs,solver_kernel.cuh:170-200=boundary terms,solver_kernel.cuh:207-260=merit shell"
R="$R,solver_kernel.cuh:286-612=riccati qp,solver_kernel.cuh:613-649=cqp setup+factor,solver_kernel.cuh:650-694=cqp (2) U rows + m"
R="$R,solver_kernel.cuh:695-724=cqp (3) forward subst,solver_kernel.cuh:725-757=cqp (4) Gram,solver_kernel.cuh:758-826=cqp (5) 4x4"
R="$R,solver_kernel.cuh:827-866=cqp (6) du + back subst,solver_kernel.cuh:867-905=cqp (7) node steps"
R="$R,solver_kernel.cuh:912-972=solve setup,solver_kernel.cuh:973-990=fill_gradient,solver_kernel.cuh:991-1020=lin setup"
R="$R,solver_kernel.cuh:1021-1031=lin Phi update,solver_kernel.cuh:1032-1056=lin Gamma update,solver_kernel.cuh:1057-1068=lin defect"
R="$R,solver_kernel.cuh:1069-1130=lin stitch,solver_kernel.cuh:1131-1161=lin tail,solver_kernel.cuh:1162-1264=qp call+derivatives+penalty"
R="$R,solver_kernel.cuh:1265-1371=line search scalar,solver_kernel.cuh:1372-1436=update+exit,solver_kernel.cuh:1441-1531=kernel setup"
R="$R,solver_kernel.cuh:1532-1588=initial guess shell,solver_kernel.cuh:1589-1711=outputs shell"
python "$(dirname "$0")/ncu_lines.py" "$1" "${2:-cart_pole_mpc_b200/libcpmpc.so}" 'cpmpc_solve_kernelILi4E' --top "${3:-0}" --ranges "$R"
