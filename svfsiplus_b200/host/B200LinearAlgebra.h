// B200LinearAlgebra — the plug-in class that exposes the CUDA backend (include/svb200.h) to
// svMultiPhysics through the reference's own LinearAlgebra interface
// (Code/Source/solver/LinearAlgebra.h:39-63), beside FsilsLinearAlgebra / PetscLinearAlgebra /
// TrilinosLinearAlgebra.  This file is compiled INSIDE the reference's source tree (it includes the
// reference's headers); INTEGRATION.md lists the registration lines.  oracle/Makefile compiles it
// against the unmodified reference headers to prove that it builds and tests/test_dropin.py drives it
// through the reference's own ComMod / eqType / FSILS_lhsType objects.
//
// Two modes, chosen by <Linear_algebra type="b200"> <Assembly> :
//   assembly "fsils" (or none)  the reference assembles on the host (do_assem into com_mod.R / com_mod.Val);
//                               solve() uploads both, solves on the GPU, returns the solution in com_mod.R.
//   assembly "b200"             global_eq_assem calls assemble_mesh(): whole-mesh element assembly on the
//                               GPU; assemble() (boundary elements) is staged and flushed by one scatter
//                               kernel; anything host code added to com_mod.R meanwhile is added on upload.
#ifndef B200_LINEAR_ALGEBRA_H
#define B200_LINEAR_ALGEBRA_H

#include "LinearAlgebra.h"
#include "CepMod.h"

#include <map>
#include <set>
#include <string>

#include "svb200.h"

// A maintainer adds `b200` to consts::LinearAlgebraType (consts.h:503); until then the class can be
// compiled against the unmodified header by defining the enumerator value on the command line.
#ifndef B200_LINEAR_ALGEBRA_TYPE
#define B200_LINEAR_ALGEBRA_TYPE consts::LinearAlgebraType::b200
#endif

class B200LinearAlgebra : public virtual LinearAlgebra {
  public:
    B200LinearAlgebra();          // cheap, no CUDA context: Parameters.cpp:2444 instantiates it at parse time
    ~B200LinearAlgebra();

    virtual void alloc(ComMod& com_mod, eqType& lEq);
    virtual void assemble(ComMod& com_mod, const int num_elem_nodes, const Vector<int>& eqN,
        const Array3<double>& lK, const Array<double>& lR);
    virtual void check_options(const consts::PreconditionerType prec_cond_type, const consts::LinearAlgebraType assembly_type);
    virtual void initialize(ComMod& com_mod, eqType& lEq);
    virtual void solve(ComMod& com_mod, eqType& lEq, const Vector<int>& incL, const Vector<double>& res);
    virtual void set_assembly(consts::LinearAlgebraType atype);
    virtual void set_preconditioner(consts::PreconditionerType prec_type);

    /// Whole-mesh element assembly on the device; called by eq_assem::global_eq_assem instead of
    /// construct_fluid when device assembly is selected.  Returns false when this physics / element
    /// type has no device kernel yet (the caller then falls back to the reference's construct_* with
    /// per-element assemble()).
    bool assemble_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag, const Array<double>& Yg,
        const Array<double>& Dg, const CepMod* cep_mod = nullptr);

    /// ustruct_r counterpart (ustruct.cpp:1726; called from main.cpp:526 on the first Newton iteration): the
    /// displacement tangent Kd lives on the device when the ustruct equation was assembled there.  Returns false
    /// when the system was assembled on the host (the caller then runs the reference's ustruct_r).
    bool ustruct_r(ComMod& com_mod, const Array<double>& Yg);

    /// b_assem_neu_bc counterpart (eq_assem.cpp:58; called from set_bc::set_bc_neu_l, set_bc.cpp:1449): the Neumann
    /// face is assembled on the device on top of the volume assembly of this Newton iteration.  Returns false when the
    /// face / physics has no device kernel or the volume was assembled on the host (the caller then runs the
    /// reference's b_assem_neu_bc with per-element assemble()).
    bool assemble_face(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Yg);

    /// b_neu_folw_p counterpart (eq_assem.cpp:186; called from set_bc_neu_l when lBc.flwP, set_bc.cpp:1446): follower
    /// pressure load on a struct face, on top of the device volume assembly.  Same fall-back rule as assemble_face.
    bool assemble_follower_face(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Dg);

    /// fsils_bc_update counterpart: re-upload the face vectors (moving meshes, follower loads).
    void update_faces(ComMod& com_mod);

    /// Where the work ran so far (also printed by the destructor): counts of the three assembly hooks that ran on the device and
    /// that fell back to the reference's host path, and of the solves.
    struct Stats { long solves = 0; long device[3] = {0, 0, 0}; long host[3] = {0, 0, 0}; unsigned announced = 0; };
    const Stats& stats() const { return stats_; }
    void report() const;

    bool device_assembly() const { return device_assembly_; }
    void set_device(int device) { device_ = device; }

  private:
    void check(int rc, const char* what);
    void note(int which, bool on_device, const std::string& what);
    bool assemble_mesh_impl(ComMod& com_mod, const mshType& lM, const Array<double>& Ag, const Array<double>& Yg,
        const Array<double>& Dg, const CepMod* cep_mod);
    bool assemble_face_impl(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Yg);
    bool assemble_follower_face_impl(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Dg);
    Stats stats_;
    void upload_structure(ComMod& com_mod);
    void upload_mesh(ComMod& com_mod, const mshType& lM);
    bool assemble_fluid_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag, const Array<double>& Yg);
    bool assemble_fsi_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag, const Array<double>& Yg,
        const Array<double>& Dg, const CepMod* cep_mod);
    bool fill_fluid_props(ComMod& com_mod, const eqType& eq, const dmnType& dmn, b200_fluid_props& p);
    bool fill_struct_props(ComMod& com_mod, const eqType& eq, const dmnType& dmn, b200_struct_props& p);
    static void fibre_stress(const ComMod& com_mod, const fibStrsType& Tf, double& Tfa, double& Tsa);
    bool assemble_ustruct_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag, const Array<double>& Yg,
        const Array<double>& Dg, const CepMod* cep_mod);
    bool assemble_solid_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag, const Array<double>& Yg,
        const Array<double>& Dg, const CepMod* cep_mod);
    bool assemble_domains_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag, const Array<double>& Yg,
        const Array<double>& Dg, const CepMod* cep_mod);

    b200_handle* h_ = nullptr;
    int device_ = -1;         // -1: rank mod device count, chosen in initialize()
    bool device_assembly_ = false;
    bool structure_uploaded_ = false;
    const mshType* mesh_uploaded_ = nullptr;
    const mshType* domains_uploaded_ = nullptr;
    bool any_device_contribution_ = false;
    bool ustruct_on_device_ = false;
    bool prestress_on_device_ = false;    // a prestress field / pstEq was set on the handle (cleared when it disappears)
    bool do_uploaded_ = false;            // com_mod.Do is on the device (moving-mesh Neumann faces)
    std::map<const faceType*, int> face_meshes_;      // faces whose connectivity is on the device -> slot
    static std::set<consts::LinearAlgebraType> valid_assemblers;
};

#endif
