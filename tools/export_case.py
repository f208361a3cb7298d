"""Write the synthetic pipe of SURVEY.md 8(d) as a case directory in the layout of the reference's tests/cases/fluid/pipe_RCR_3d
(mesh/mesh-complete.mesh.vtu, mesh/mesh-surfaces/lumen_{inlet,outlet,wall}.vtp, solver.xml, lumen_inlet.flow), with the VTK-free
writer (svfsiplus_b200/sv_io.py -> libsvb200io.so).  A maintainer who has the real `svmultiphysics` binary can then run the SAME
workload the bench times (P10: --dims 96 96 181) through the reference with MPI and through the B200 backend:

    python tools/export_case.py --dims 96 96 181 --out /tmp/pipe_p10        # ~10 M tets, a few hundred MB
    cd /tmp/pipe_p10 && mpiexec -n 16 svmultiphysics solver.xml              # reference, FSILS
    sed -i 's/type="fsils"/type="b200"/' solver.xml && svmultiphysics solver.xml   # this backend

Node and element ids are 1-based in the files (GlobalNodeID point data, GlobalElementID cell data), as the reference's loaders
expect (vtk_xml_parser.cpp: faces subtract one).  The solver parameters are those of the reference case (rho 1.06, mu 0.04, dt 0.005,
rho_inf 0.5, LS NS 1e-3 / GM 1e-3 x 10 / CG 1e-3 x 300, Krylov 250, RCR outlet C 1.5e-5, Rd 1212, Rp 121, backflow 0.2); the inflow
is a smooth synthetic pulse (one period, 33 samples, 16 Fourier modes), not the reference's measured waveform.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svfsiplus_b200 import mesh as M      # noqa: E402
from svfsiplus_b200 import sv_io as IO    # noqa: E402

FACES = {"lumen_inlet": "inlet", "lumen_outlet": "outlet", "lumen_wall": "wall"}

SOLVER_XML = """<?xml version="1.0" encoding="UTF-8" ?>
<svMultiPhysicsFile version="0.1">
<GeneralSimulationParameters>
  <Continue_previous_simulation> false </Continue_previous_simulation>
  <Number_of_spatial_dimensions> 3 </Number_of_spatial_dimensions>
  <Number_of_time_steps> {steps} </Number_of_time_steps>
  <Time_step_size> 0.005 </Time_step_size>
  <Spectral_radius_of_infinite_time_step> 0.50 </Spectral_radius_of_infinite_time_step>
  <Searched_file_name_to_trigger_stop> STOP_SIM </Searched_file_name_to_trigger_stop>
  <Save_results_to_VTK_format> 1 </Save_results_to_VTK_format>
  <Name_prefix_of_saved_VTK_files> result </Name_prefix_of_saved_VTK_files>
  <Increment_in_saving_VTK_files> {steps} </Increment_in_saving_VTK_files>
  <Start_saving_after_time_step> 1 </Start_saving_after_time_step>
  <Increment_in_saving_restart_files> 100 </Increment_in_saving_restart_files>
  <Convert_BIN_to_VTK_format> 0 </Convert_BIN_to_VTK_format>
  <Verbose> 1 </Verbose>
  <Warning> 0 </Warning>
  <Debug> 0 </Debug>
</GeneralSimulationParameters>
<Add_mesh name="msh" >
  <Mesh_file_path> mesh/mesh-complete.mesh.vtu </Mesh_file_path>
{faces}</Add_mesh>
<Add_equation type="fluid" >
  <Coupled> true </Coupled>
  <Min_iterations> 3 </Min_iterations>
  <Max_iterations> 5 </Max_iterations>
  <Tolerance> 1e-11 </Tolerance>
  <Backflow_stabilization_coefficient> 0.2 </Backflow_stabilization_coefficient>
  <Density> 1.06 </Density>
  <Viscosity model="Constant" > <Value> 0.04 </Value> </Viscosity>
  <Output type="Spatial" > <Velocity> true </Velocity> <Pressure> true </Pressure> </Output>
  <LS type="NS" >
    {linear_algebra}
    <Max_iterations> 15 </Max_iterations>
    <NS_GM_max_iterations> 10 </NS_GM_max_iterations>
    <NS_CG_max_iterations> 300 </NS_CG_max_iterations>
    <Tolerance> 1e-3 </Tolerance>
    <NS_GM_tolerance> 1e-3 </NS_GM_tolerance>
    <NS_CG_tolerance> 1e-3 </NS_CG_tolerance>
    <Absolute_tolerance> 1e-17 </Absolute_tolerance>
    <Krylov_space_dimension> 250 </Krylov_space_dimension>
  </LS>
  <Add_BC name="lumen_inlet" >
    <Type> Dir </Type> <Time_dependence> Unsteady </Time_dependence>
    <Temporal_values_file_path> lumen_inlet.flow </Temporal_values_file_path>
    <Profile> Parabolic </Profile> <Impose_flux> true </Impose_flux>
  </Add_BC>
  <Add_BC name="lumen_outlet" >
    <Type> Neu </Type> <Time_dependence> RCR </Time_dependence>
    <RCR_values>
      <Capacitance> 1.5e-5 </Capacitance> <Distal_resistance> 1212 </Distal_resistance> <Proximal_resistance> 121 </Proximal_resistance>
      <Distal_pressure> 0 </Distal_pressure> <Initial_pressure> 0 </Initial_pressure>
    </RCR_values>
  </Add_BC>
  <Add_BC name="lumen_wall" > <Type> Dir </Type> <Time_dependence> Steady </Time_dependence> <Value> 0.0 </Value> </Add_BC>
</Add_equation>
</svMultiPhysicsFile>
"""


LINEAR_ALGEBRA = {
    # the reference's own backend
    "fsils": '<Linear_algebra type="fsils" > <Preconditioner> fsils </Preconditioner> </Linear_algebra>',
    # this repository's backend: whole-mesh assembly + solve on the GPU / host assembly + GPU solve
    "b200": '<Linear_algebra type="b200" > <Preconditioner> fsils </Preconditioner> <Assembly> b200 </Assembly> </Linear_algebra>',
    "b200_solve_only": '<Linear_algebra type="b200" > <Preconditioner> fsils </Preconditioner> </Linear_algebra>',
}


def export_pipe(out, dims, steps=2, mode=IO.APPENDED_RAW, linear_algebra="fsils"):
    """Returns dict(nNo, nEl, faces={name: (n_nodes, n_elems)})."""
    nx, ny, nz = dims
    m = M.pipe_mesh(nx, ny, nz)
    os.makedirs(os.path.join(out, "mesh", "mesh-surfaces"), exist_ok=True)
    IO.write_vtk(os.path.join(out, "mesh", "mesh-complete.mesh.vtu"), m.x, m.ien, IO.VTK_TYPE["TET4"],
                 {"GlobalNodeID": np.arange(1, m.nNo + 1, dtype=np.int32)}, {"GlobalElementID": np.arange(1, m.nEl + 1, dtype=np.int32)}, mode=mode)
    info = dict(nNo=m.nNo, nEl=m.nEl, faces={})
    xml_faces = ""
    for name, key in FACES.items():
        nodes = np.asarray(m.faces[key]["nodes"])
        on = np.zeros(m.nNo, bool)
        on[nodes] = True
        IENb, gE = M.face_elements(m, on)
        loc = np.full(m.nNo, -1, np.int64)
        loc[nodes] = np.arange(len(nodes))
        IO.write_vtk(os.path.join(out, "mesh", "mesh-surfaces", name + ".vtp"), m.x[nodes], loc[IENb].astype(np.int32), IO.VTK_TYPE["TRI3"],
                     {"GlobalNodeID": (nodes + 1).astype(np.int32)}, {"GlobalElementID": (gE + 1).astype(np.int32)}, polydata=True, mode=mode)
        info["faces"][name] = (len(nodes), len(gE))
        xml_faces += f'  <Add_face name="{name}"> <Face_file_path> mesh/mesh-surfaces/{name}.vtp </Face_file_path> </Add_face>\n'
    with open(os.path.join(out, "solver.xml"), "w") as f:
        f.write(SOLVER_XML.format(steps=steps, faces=xml_faces, linear_algebra=LINEAR_ALGEBRA[linear_algebra]))
    # one smooth pulse per second, peak inflow 50 mL/s (negative = into the domain, like the reference case)
    t = np.linspace(0.0, 1.0, 33)
    q = -50.0 * np.sin(np.pi * t) ** 2
    with open(os.path.join(out, "lumen_inlet.flow"), "w") as f:
        f.write("33    16\n")
        for ti, qi in zip(t, q):
            f.write(f"{ti:.6f}    {qi:.6f}\n")
    return info


BLOCK_XML = """<?xml version="1.0" encoding="UTF-8" ?>
<svMultiPhysicsFile version="0.1">
<GeneralSimulationParameters>
  <Continue_previous_simulation> false </Continue_previous_simulation>
  <Number_of_spatial_dimensions> 3 </Number_of_spatial_dimensions>
  <Number_of_time_steps> {steps} </Number_of_time_steps>
  <Time_step_size> 1e-4 </Time_step_size>
  <Spectral_radius_of_infinite_time_step> 0.50 </Spectral_radius_of_infinite_time_step>
  <Save_results_to_VTK_format> 1 </Save_results_to_VTK_format>
  <Name_prefix_of_saved_VTK_files> result </Name_prefix_of_saved_VTK_files>
  <Increment_in_saving_VTK_files> {steps} </Increment_in_saving_VTK_files>
  <Start_saving_after_time_step> 1 </Start_saving_after_time_step>
  <Increment_in_saving_restart_files> 100 </Increment_in_saving_restart_files>
  <Verbose> 1 </Verbose>
</GeneralSimulationParameters>
<Add_mesh name="msh" >
  <Mesh_file_path> mesh/mesh-complete.mesh.vtu </Mesh_file_path>
{faces}</Add_mesh>
<Add_equation type="struct" >
  <Coupled> true </Coupled>
  <Min_iterations> 1 </Min_iterations>
  <Max_iterations> 3 </Max_iterations>
  <Tolerance> 1e-9 </Tolerance>
  <Constitutive_model type="nHK"> </Constitutive_model>
  <Density> 1000.0 </Density>
  <Elasticity_modulus> 240.56596E6 </Elasticity_modulus>
  <Poisson_ratio> 0.5 </Poisson_ratio>
  <Dilational_penalty_model> ST91 </Dilational_penalty_model>
  <Penalty_parameter> 4.0E9 </Penalty_parameter>
  <Output type="Spatial" > <Displacement> true </Displacement> <Velocity> true </Velocity> </Output>
  <LS type="BICG" >
    {linear_algebra}
    <Tolerance> 1e-12 </Tolerance>
    <Max_iterations> 600 </Max_iterations>
  </LS>
  <Add_BC name="X0" > <Type> Dir </Type> <Value> 0.0 </Value> <Effective_direction> (0, 0, 1) </Effective_direction> </Add_BC>
  <Add_BC name="Y0" > <Type> Dir </Type> <Value> 0.0 </Value> <Effective_direction> (0, 1, 0) </Effective_direction> </Add_BC>
  <Add_BC name="Z0" > <Type> Dir </Type> <Value> 0.0 </Value> <Effective_direction> (1, 0, 0) </Effective_direction> </Add_BC>
  <Add_BC name="Z1" > <Type> Neu </Type> <Time_dependence> Steady </Time_dependence> <Value> 5.0e6 </Value> </Add_BC>
</Add_equation>
</svMultiPhysicsFile>
"""


USTRUCT_XML = """<?xml version="1.0" encoding="UTF-8" ?>
<svMultiPhysicsFile version="0.1">
<GeneralSimulationParameters>
  <Continue_previous_simulation> 0 </Continue_previous_simulation>
  <Number_of_spatial_dimensions> 3 </Number_of_spatial_dimensions>
  <Number_of_time_steps> {steps} </Number_of_time_steps>
  <Time_step_size> 0.01 </Time_step_size>
  <Spectral_radius_of_infinite_time_step> 0.50 </Spectral_radius_of_infinite_time_step>
  <Searched_file_name_to_trigger_stop> STOP_SIM </Searched_file_name_to_trigger_stop>
  <Save_results_to_VTK_format> 1 </Save_results_to_VTK_format>
  <Name_prefix_of_saved_VTK_files> result </Name_prefix_of_saved_VTK_files>
  <Increment_in_saving_VTK_files> {steps} </Increment_in_saving_VTK_files>
  <Start_saving_after_time_step> 1 </Start_saving_after_time_step>
  <Increment_in_saving_restart_files> 100 </Increment_in_saving_restart_files>
  <Convert_BIN_to_VTK_format> 0 </Convert_BIN_to_VTK_format>
  <Verbose> 1 </Verbose>
  <Warning> 0 </Warning>
  <Debug> 0 </Debug>
</GeneralSimulationParameters>
<Add_mesh name="msh" >
  <Mesh_file_path> mesh/mesh-complete.mesh.vtu </Mesh_file_path>
{faces}</Add_mesh>
<Add_equation type="ustruct" >
  <Coupled> true </Coupled>
  <Min_iterations> 3 </Min_iterations>
  <Max_iterations> 5 </Max_iterations>
  <Tolerance> 1e-12 </Tolerance>
  <Constitutive_model type="nHK"> </Constitutive_model>
  <Density> 1e-3 </Density>
  <Elasticity_modulus> 240.56596e6 </Elasticity_modulus>
  <Poisson_ratio> 0.4999999 </Poisson_ratio>
  <Dilational_penalty_model> ST91 </Dilational_penalty_model>
  <Momentum_stabilization_coefficient> 1e-3 </Momentum_stabilization_coefficient>
  <Continuity_stabilization_coefficient> 1e-3 </Continuity_stabilization_coefficient>
  <Output type="Spatial" > <Displacement> true </Displacement> <Velocity> true </Velocity> <Pressure> true </Pressure> </Output>
  <LS type="GMRES" >
    {linear_algebra}
    <Tolerance> {ls_tol} </Tolerance>
    <Max_iterations> 100 </Max_iterations>
    <Krylov_space_dimension> 300 </Krylov_space_dimension>
  </LS>
  <Add_BC name="X0" > <Type> Dir </Type> <Value> 0.0 </Value> <Effective_direction> (1, 0, 0) </Effective_direction>
    <Impose_on_state_variable_integral> true </Impose_on_state_variable_integral> </Add_BC>
  <Add_BC name="Y0" > <Type> Dir </Type> <Value> 0.0 </Value> <Effective_direction> (0, 1, 0) </Effective_direction>
    <Impose_on_state_variable_integral> true </Impose_on_state_variable_integral> </Add_BC>
  <Add_BC name="Z0" > <Type> Dir </Type> <Value> 0.0 </Value> <Effective_direction> (0, 0, 1) </Effective_direction>
    <Impose_on_state_variable_integral> true </Impose_on_state_variable_integral> </Add_BC>
  <Add_BC name="Z1" > <Type> Neu </Type> <Time_dependence> Steady </Time_dependence> <Value> 5.0e6 </Value>
    <Follower_pressure_load> {follower} </Follower_pressure_load> </Add_BC>
</Add_equation>
</svMultiPhysicsFile>
"""


def export_block(out, n, elem="hex", steps=1, mode=IO.APPENDED_RAW, linear_algebra="fsils", phys="struct", follower=True, ls_tol="1e-6"):
    """The solid block of SURVEY 8(d) (n^3 HEX8, its 6-tet split, or the quadratic split) in the layout of the reference's
    tests/cases/struct/block_compression: volume mesh, the six faces X0..Z1 (QUD4 / TRI3 / TRI6), a struct solver.xml."""
    m = M.block_mesh(n, elem)
    os.makedirs(os.path.join(out, "mesh", "mesh-surfaces"), exist_ok=True)
    vt = {"hex": "HEX8", "tet": "TET4", "tet10": "TET10"}[elem]
    ft = {"hex": "QUD4", "tet": "TRI3", "tet10": "TRI6"}[elem]
    IO.write_vtk(os.path.join(out, "mesh", "mesh-complete.mesh.vtu"), m.x, m.ien, IO.VTK_TYPE[vt],
                 {"GlobalNodeID": np.arange(1, m.nNo + 1, dtype=np.int32)}, {"GlobalElementID": np.arange(1, m.nEl + 1, dtype=np.int32)}, mode=mode)
    info = dict(nNo=m.nNo, nEl=m.nEl, eNoN=m.ien.shape[1], faces={})
    xml_faces = ""
    for name in ("X0", "X1", "Y0", "Y1", "Z0", "Z1"):
        nodes = np.asarray(m.faces[name]["nodes"])
        on = np.zeros(m.nNo, bool)
        on[nodes] = True
        IENb, gE = M.face_elements(m, on)
        loc = np.full(m.nNo, -1, np.int64)
        loc[nodes] = np.arange(len(nodes))
        IO.write_vtk(os.path.join(out, "mesh", "mesh-surfaces", name + ".vtp"), m.x[nodes], loc[IENb].astype(np.int32), IO.VTK_TYPE[ft],
                     {"GlobalNodeID": (nodes + 1).astype(np.int32)}, {"GlobalElementID": (gE + 1).astype(np.int32)}, polydata=True, mode=mode)
        info["faces"][name] = (len(nodes), len(gE), IENb.shape[1])
        xml_faces += f'  <Add_face name="{name}"> <Face_file_path> mesh/mesh-surfaces/{name}.vtp </Face_file_path> </Add_face>\n'
    with open(os.path.join(out, "solver.xml"), "w") as f:
        if phys == "ustruct":        # tests/cases/ustruct/block_compression/P1P1_VMS/solver.xml (steady load instead of the ramp file)
            f.write(USTRUCT_XML.format(steps=steps, faces=xml_faces, linear_algebra=LINEAR_ALGEBRA[linear_algebra],
                                       follower="true" if follower else "false", ls_tol=ls_tol))
        else:
            f.write(BLOCK_XML.format(steps=steps, faces=xml_faces, linear_algebra=LINEAR_ALGEBRA[linear_algebra]))
    return info


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dims", type=int, nargs=3, default=[24, 24, 48], help="pipe hex counts nx ny nz (P10 = 96 96 181)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--block", type=int, default=0, help="write the solid block with this many elements per edge instead of the pipe")
    ap.add_argument("--elem", default="hex", choices=["hex", "tet", "tet10"])
    ap.add_argument("--linear-algebra", default="fsils", choices=sorted(LINEAR_ALGEBRA))
    a = ap.parse_args()
    print(export_block(a.out, a.block, a.elem, a.steps, linear_algebra=a.linear_algebra) if a.block
          else export_pipe(a.out, tuple(a.dims), a.steps, linear_algebra=a.linear_algebra))
