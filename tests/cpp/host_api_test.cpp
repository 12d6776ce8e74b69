// Host-only unit test of the C++ mirror of ParElag's solver-side API (no device needed): the pieces a C++ driver of the
// reference touches besides BuildSolver / Mult -- ParameterList + XML reader, the SolverLibrary registry, Level,
// SolverState, TimeManager, MfemBlockOperator offsets, Hierarchy::SetCycle.  Built and run by tests/test_host_api_cpu.py.
#include "parelag_solvers.hpp"
#include <cstdio>
#include <set>

using namespace parelag;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)
template <class E, class F> static bool throws(F f) { try { f(); } catch (const E &) { return true; } catch (...) { return false; } return false; }

// a user-defined factory type, as in the example of ParELAG_SolverLibrary.hpp:46-58
struct IdentitySolver : Solver
{
    explicit IdentitySolver(int n) : Solver(n, n, false) {}
    void Mult(const mfem::Vector &, mfem::Vector &) const override {}
    void MultTranspose(const mfem::Vector &, mfem::Vector &) const override {}
    void _do_set_operator(const std::shared_ptr<mfem::Operator> &) override {}
};
struct SuperCoolSolverFactory : SolverFactory
{
    std::unique_ptr<mfem::Solver> _do_build_solver(const std::shared_ptr<mfem::Operator> &op, SolverState &) const override
    { return std::make_unique<IdentitySolver>(op->Height()); }
    void _do_set_default_parameters() override { GetParameters().Get("Coolness", 11); }
    void _do_initialize(const ParameterList &) override {}
};
struct DummyOp : mfem::Operator
{
    explicit DummyOp(int n) : mfem::Operator(n) {}
    void Mult(const mfem::Vector &, mfem::Vector &) const override {}
};

int main()
{
    // ---- ParameterList: Get with a default inserts it; typed access; sublists; Merge
    {
        ParameterList pl("top");
        CHECK(pl.Get("tol", 1e-6) == 1e-6 && pl.IsParameter("tol"));
        CHECK(throws<bad_var_cast>([&] { (void)pl.Get<int>("tol"); }));
        CHECK(throws<std::out_of_range>([&] { (void)pl.Get<int>("absent"); }));
        pl.Sublist("child").Set("n", 3);
        ParameterList other("o");
        other.Sublist("child").Set("m", 4);
        other.Set("tol", 1e-3);
        pl.Merge(other);
        CHECK(pl.Sublist("child").Get<int>("n") == 3 && pl.Sublist("child").Get<int>("m") == 4 && pl.Get<double>("tol") == 1e-3);
        CHECK(throws<std::out_of_range>([&] { (void)pl.Sublist("nope", true); }));
    }
    // ---- XML -> library -> registry
    {
        const char *xml =
            "<ParameterList name=\"Preconditioner Library\">"
            "  <ParameterList name=\"GS\"><Parameter name=\"Type\" type=\"string\" value=\"Hypre\"/>"
            "    <ParameterList name=\"Solver Parameters\"><Parameter name=\"Type\" type=\"string\" value=\"L1 Gauss-Seidel\"/></ParameterList></ParameterList>"
            "  <ParameterList name=\"Cool\"><Parameter name = \"Type\" type=\"string\" value=\"SuperCool\"/></ParameterList>"
            "</ParameterList>";
        SimpleXMLParameterListReader reader;
        auto pl = reader.Parse(xml);
        auto lib = SolverLibrary::CreateLibrary(*pl);
        auto names = lib->GetSolverNames();
        CHECK((std::set<std::string>(names.begin(), names.end()) == std::set<std::string>{"GS", "Cool"}));
        CHECK(lib->IsSolver("GS") && !lib->IsSolver("nope"));
        auto types = lib->GetSolverFactoryNames();
        const std::set<std::string> tset(types.begin(), types.end());
        for (const char *t : {"AMGe", "Hypre", "Hiptmair", "Krylov", "Stationary Iteration", "Block GS", "Block Jacobi", "Block LDU"}) CHECK(tset.count(t) == 1);
        CHECK(throws<std::runtime_error>([&] { (void)lib->GetSolverFactory("Cool"); }));         // unknown type "SuperCool"
        CHECK(lib->AddNewSolverFactory("SuperCool", [] { return std::make_shared<SuperCoolSolverFactory>(); }));
        CHECK(!lib->AddNewSolverFactory("SuperCool", [] { return std::make_shared<SuperCoolSolverFactory>(); }));   // already there
        auto fact = lib->GetSolverFactory("Cool");
        CHECK(fact && fact->GetParameters().Get<int>("Coolness") == 11);
        auto op = std::make_shared<DummyOp>(7);
        auto state = fact->GetDefaultState();
        std::shared_ptr<mfem::Operator> base = op;
        auto solver = fact->BuildSolver(base, *state);
        CHECK(solver && solver->Height() == 7);
        CHECK(lib->RemoveSolverFactory("SuperCool") && !lib->RemoveSolverFactory("SuperCool"));
        CHECK(throws<std::out_of_range>([&] { (void)lib->GetSolverFactory("nope"); }));
        CHECK(fact->GetDefaultState() != nullptr);
    }
    // ---- the scenario of the reference's own unit test of the library (src/linalg/unit_test/solver_lib_test.cpp): a library
    //      filled from a ParameterList built in code, a Hiptmair factory created by hand and wired to the library, the
    //      nested factories it resolves
    {
        ParameterList master("master");
        master.Sublist("Solver1").Set("Type", "Hypre");
        master.Sublist("Solver1").Sublist("Solver Parameters");
        master.Sublist("Solver2").Set("Type", "Hiptmair");
        master.Sublist("Solver2").Sublist("Solver Parameters").Set("Primary Smoother", "Solver1");
        master.Sublist("Solver2").Sublist("Solver Parameters").Set("Auxiliary Smoother", "Solver1");
        auto lib = SolverLibrary::CreateLibrary();
        lib->Initialize(master);
        auto names = lib->GetSolverNames();
        CHECK((std::set<std::string>(names.begin(), names.end()) == std::set<std::string>{"Solver1", "Solver2"}));
        auto hand_made = std::make_shared<HiptmairSmootherFactory>();
        hand_made->SetSolverLibrary(lib);
        ParameterList wiring;
        wiring.Set("Primary Smoother", "Solver1");
        wiring.Set("Auxiliary Smoother", "Solver1");
        hand_made->Initialize(wiring);
        CHECK(std::dynamic_pointer_cast<HypreSmootherFactory>(hand_made->GetPrimarySmootherFactory()) != nullptr);
        CHECK(std::dynamic_pointer_cast<HypreSmootherFactory>(hand_made->GetAuxiliarySmootherFactory()) != nullptr);
        CHECK(hand_made->GetPrimarySmootherFactory() == lib->GetSolverFactory("Solver1"));          // factories are shared
        CHECK(std::dynamic_pointer_cast<HiptmairSmootherFactory>(lib->GetSolverFactory("Solver2")) != nullptr);
        // the defaults name a "Default Hypre" entry (ParELAG_HiptmairSmootherFactory.cpp:26-31), absent from this library
        auto defaults = std::make_shared<HiptmairSmootherFactory>();
        defaults->SetSolverLibrary(lib);
        CHECK(throws<std::out_of_range>([&] { defaults->Initialize(ParameterList()); }));
        CHECK(defaults->GetParameters().Get<std::string>("Primary Smoother") == "Default Hypre");
        // factories handed over directly
        HiptmairSmootherFactory direct(lib->GetSolverFactory("Solver1"));
        CHECK(direct.GetPrimarySmootherFactory() == direct.GetAuxiliarySmootherFactory());
    }
    // ---- Level
    {
        auto fine = std::make_shared<Level>(0), coarse = std::make_shared<Level>(1);
        coarse->SetPreviousLevel(fine);
        CHECK(coarse->GetPreviousLevel() == fine && coarse->GetLevelID() == 1);
        fine.reset();
        CHECK(coarse->GetPreviousLevel() == nullptr);                               // a weak reference, as in the reference
        coarse->Set<int>("n", 5);
        CHECK(coarse->IsKey("n") && coarse->Get<int>("n") == 5);
        coarse->Reset<int>("n", 6);
        CHECK(coarse->Get<int>("n") == 6);
        CHECK(throws<std::out_of_range>([&] { (void)coarse->Get<int>("absent"); }));
        CHECK(throws<bad_var_cast>([&] { (void)coarse->Get<double>("n"); }));
        coarse->Set<std::shared_ptr<mfem::Operator>>("PreSmoother", nullptr);
        CHECK(coarse->IsKey("PreSmoother") && !coarse->IsValidKey("PreSmoother"));   // a null smoother is "not valid"
    }
    // ---- SolverState, TimeManager
    {
        SolverState st;
        CHECK(!st.HasDeRhamSequence() && !st.IsVector("elemMatrixScaling") && !st.IsOperator("A"));
        st.SetForms({2, 3});
        CHECK(st.GetForms().size() == 2);
        TimeManager::ClearAllTimers();
        { Timer t = TimeManager::AddTimer("unit test timer"); }
        CHECK(TimeManager::IsTimer("unit test timer") && !TimeManager::IsTimer("other"));
        CHECK(TimeManager::DeleteTimer("unit test timer") && !TimeManager::DeleteTimer("unit test timer"));
    }
    // ---- MfemBlockOperator offsets, Hierarchy::SetCycle
    {
        MfemBlockOperator B(std::vector<MfemBlockOperator::offset_type>{0, 4, 6});
        mfem::Array<MfemBlockOperator::offset_type> ro, co;
        B.CopyRowOffsetsAsMfemArray(ro); B.CopyColumnOffsetsAsMfemArray(co);
        CHECK(ro.Size() == 3 && ro[1] == 4 && co[2] == 6 && B.GetNumBlockRows() == 2 && B.IsZeroBlock(0, 1));
        Hierarchy H(std::make_shared<DummyOp>(10), 3);
        CHECK(H.GetNumLevels() == 3);
        H.SetCycle(1, 1);
        H.SetCycle(std::vector<int>{0, 1, 0});
        H.SetCycle(2);
        CHECK(throws<std::runtime_error>([&] { H.SetCycle(-1); }));
        CHECK(throws<std::runtime_error>([&] { H.SetCycle(1, 7); }));
        CHECK(throws<std::runtime_error>([&] { H.SetCycle(std::vector<int>{1, 1}); }));
    }
    if (failures == 0) std::printf("HOST_API_TEST_OK\n");
    return failures == 0 ? 0 : 1;
}
