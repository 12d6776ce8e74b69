// parelag_core.hpp -- host-side mirror of ParElag's solver API (pure C++, no CUDA):
//   ParameterList (+ XML reader)   src/utilities/ParELAG_ParameterList.hpp:40-320,
//                                  src/utilities/ParELAG_SimpleXMLParameterListReader.cpp:220-295
//   TimeManager / Timer            src/utilities/ParELAG_TimeManager.hpp:40-146, ParELAG_Timer.hpp:30-60
//   error macros                   src/utilities/elagError.hpp:61-140
//   Level                          src/linalg/solver_core/ParELAG_Level.hpp:39-214
//   Solver                         src/linalg/solver_core/ParELAG_Solver.hpp:36-105
//   SolverState / NestedSolverState src/linalg/solver_core/ParELAG_SolverState.hpp:54-317
//   SolverFactory                  src/linalg/solver_core/ParELAG_SolverFactory.hpp:36-186
//   SolverLibrary                  src/linalg/solver_core/ParELAG_SolverLibrary.hpp:65-275
// Same names, argument meaning and error behaviour (exceptions thrown by
// PARELAG_TEST_FOR_EXCEPTION / PARELAG_ASSERT) as the reference, so the reference's
// drivers and tests read the same against this library.
#pragma once
#include <algorithm>
#include <any>
#include <cctype>
#include <chrono>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <unordered_map>
#include <vector>
#include "mfem_compat.hpp"

namespace parelag
{
// ---------------------------------------------------------------- errors
struct not_implemented_error : std::logic_error { using std::logic_error::logic_error; };
struct bad_var_cast : std::runtime_error { using std::runtime_error::runtime_error; };

#define PARELAG_TEST_FOR_EXCEPTION(cond, exc, msg)                                       \
    do {                                                                                 \
        if (cond) {                                                                      \
            std::ostringstream pe_os_;                                                   \
            pe_os_ << __FILE__ << ":" << __LINE__ << ": " << #cond << "\n" << msg;       \
            throw exc(pe_os_.str());                                                     \
        }                                                                                \
    } while (0)
#define PARELAG_ASSERT(cond) PARELAG_TEST_FOR_EXCEPTION(!(cond), std::runtime_error, "Assertion failed.")
#define PARELAG_NOT_IMPLEMENTED() throw ::parelag::not_implemented_error(std::string(__func__) + " is not implemented")

using std::make_unique;   // src/utilities/MemoryUtils.hpp provides parelag::make_unique for C++11

// ---------------------------------------------------------------- timers
class TimeManager
{
public:
    struct Watch { double total = 0.0; int count = 0; bool running = false;
                   std::chrono::steady_clock::time_point t0; };
    static std::map<std::string, Watch> &Map() { static std::map<std::string, Watch> m; return m; }
    static void Start(const std::string &n)
    {
        auto &w = Map()[n];
        if (Device::Ctx()) pe_ctx_sync(Device::Ctx());
        w.running = true; w.t0 = std::chrono::steady_clock::now();
    }
    static void Stop(const std::string &n)
    {
        auto &w = Map()[n];
        if (!w.running) return;
        if (Device::Ctx()) pe_ctx_sync(Device::Ctx());
        w.total += std::chrono::duration<double>(std::chrono::steady_clock::now() - w.t0).count();
        w.count++; w.running = false;
    }
    static double Seconds(const std::string &n) { auto it = Map().find(n); return it == Map().end() ? 0.0 : it->second.total; }
    static void Print(std::ostream &os = std::cout)
    {
        os << "Timer name : seconds (calls)\n";
        for (auto &kv : Map()) os << kv.first << " : " << kv.second.total << " (" << kv.second.count << ")\n";
    }
    static void ClearAllData() { Map().clear(); }
    /// IsTimer / DeleteTimer / ClearAllTimers / PrintSerial (ParELAG_TimeManager.hpp:52-135)
    static bool IsTimer(const std::string &n) { return Map().count(n) > 0; }
    static bool DeleteTimer(const std::string &n) { return Map().erase(n) > 0; }
    static void ClearAllTimers() { Map().clear(); }
    static void PrintSerial(std::ostream &os = std::cout) { Print(os); }
    class TimerT;
    static TimerT AddTimer(const std::string &n);
    static TimerT GetTimer(const std::string &n);
};
/// RAII timer: starts in the constructor, stops in the destructor or at Stop()
class TimeManager::TimerT
{
public:
    explicit TimerT(std::string n) : name_(std::move(n)) { TimeManager::Start(name_); }
    TimerT(TimerT &&o) noexcept : name_(std::move(o.name_)), stopped_(o.stopped_) { o.stopped_ = true; }
    ~TimerT() { Stop(); }
    void Stop() { if (!stopped_) { TimeManager::Stop(name_); stopped_ = true; } }
private:
    std::string name_;
    bool stopped_ = false;
};
using Timer = TimeManager::TimerT;
inline Timer TimeManager::AddTimer(const std::string &n) { return Timer(n); }
inline Timer TimeManager::GetTimer(const std::string &n) { return Timer(n); }

// ---------------------------------------------------------------- ParameterList
class ParameterList
{
public:
    using key_type = std::string;
    ParameterList() = default;
    explicit ParameterList(std::string name) : Name_(std::move(name)) {}
    ParameterList(const ParameterList &rhs) { *this = rhs; }
    ParameterList &operator=(const ParameterList &rhs)
    {
        if (this == &rhs) return *this;
        Name_ = rhs.Name_; Params_ = rhs.Params_; Sublists_.clear();
        for (auto &kv : rhs.Sublists_) Sublists_[kv.first] = make_unique<ParameterList>(*kv.second);
        return *this;
    }
    ParameterList(ParameterList &&) = default;
    ParameterList &operator=(ParameterList &&) = default;
    virtual ~ParameterList() = default;

    const std::string &GetName() const noexcept { return Name_; }
    void SetName(std::string name) { Name_ = std::move(name); }
    bool IsParameter(const key_type &name) const noexcept { return Params_.count(name) > 0; }
    bool IsSublist(const key_type &name) const noexcept { return Sublists_.count(name) > 0; }
    bool IsValid(const key_type &name) const noexcept { return IsParameter(name) || IsSublist(name); }

    template <typename T>
    void Set(const key_type &name, T &&val) { Params_[name] = std::any(typename std::decay<T>::type(std::forward<T>(val))); }
    void Set(const key_type &name, const char *val) { Params_[name] = std::any(std::string(val)); }
    void Set(const key_type &name, const ParameterList &val) { Sublists_[name] = make_unique<ParameterList>(val); }

    /// Get with default: INSERTS the default when absent (reference behaviour,
    /// ParELAG_ParameterList.hpp:172-189)
    template <typename T>
    T &Get(const key_type &name, const T &default_value)
    {
        auto it = Params_.find(name);
        if (it == Params_.end()) it = Params_.emplace(name, std::any(default_value)).first;
        return Cast<T>(it->second, name);
    }
    std::string &Get(const key_type &name, const char default_value[]) { return Get<std::string>(name, std::string(default_value)); }
    template <typename T>
    T &Get(const key_type &name)
    {
        auto it = Params_.find(name);
        PARELAG_TEST_FOR_EXCEPTION(it == Params_.end(), std::out_of_range,
                                   "ParameterList::Get(): parameter \"" << name << "\" not found in list \"" << Name_ << "\".");
        return Cast<T>(it->second, name);
    }
    template <typename T>
    const T &Get(const key_type &name) const { return const_cast<ParameterList *>(this)->Get<T>(name); }

    void Merge(const ParameterList &other)
    {
        for (auto &kv : other.Params_) Params_[kv.first] = kv.second;
        for (auto &kv : other.Sublists_)
        {
            if (IsSublist(kv.first)) Sublists_[kv.first]->Merge(*kv.second);
            else Sublists_[kv.first] = make_unique<ParameterList>(*kv.second);
        }
    }
    ParameterList &Sublist(const std::string &name, bool must_exist = false)
    {
        auto it = Sublists_.find(name);
        if (it == Sublists_.end())
        {
            PARELAG_TEST_FOR_EXCEPTION(must_exist, std::out_of_range,
                                       "ParameterList::Sublist(): sublist \"" << name << "\" does not exist.");
            it = Sublists_.emplace(name, make_unique<ParameterList>(name)).first;
        }
        return *it->second;
    }
    const ParameterList &Sublist(const std::string &name) const
    {
        auto it = Sublists_.find(name);
        PARELAG_TEST_FOR_EXCEPTION(it == Sublists_.end(), std::out_of_range,
                                   "ParameterList::Sublist(): sublist \"" << name << "\" does not exist.");
        return *it->second;
    }
    std::vector<std::string> SublistNames() const
    {
        std::vector<std::string> out;
        for (auto &kv : Sublists_) out.push_back(kv.first);
        return out;
    }
    /// every parameter as "path/name<TAB>type<TAB>value", sublists separated by '/', lines in lexicographic order;
    /// integers in decimal, floating point with 17 significant digits, bool true/false, vectors blank separated,
    /// string lists comma separated (introspection for tests of the XML reader)
    void Dump(std::vector<std::string> &out, const std::string &prefix = "") const
    {
        for (auto &kv : Params_)
        {
            std::ostringstream v;
            v.precision(17);
            std::string type = "?";
            const std::any &a = kv.second;
            if (auto p = std::any_cast<bool>(&a)) { type = "bool"; v << (*p ? "true" : "false"); }
            else if (auto p = std::any_cast<char>(&a)) { type = "char"; v << (int)*p; }
            else if (auto p = std::any_cast<int>(&a)) { type = "int"; v << *p; }
            else if (auto p = std::any_cast<long>(&a)) { type = "long"; v << *p; }
            else if (auto p = std::any_cast<unsigned long>(&a)) { type = "unsigned long"; v << *p; }
            else if (auto p = std::any_cast<long long>(&a)) { type = "long long"; v << *p; }
            else if (auto p = std::any_cast<unsigned long long>(&a)) { type = "unsigned long long"; v << *p; }
            else if (auto p = std::any_cast<float>(&a)) { type = "float"; v << (double)*p; }
            else if (auto p = std::any_cast<double>(&a)) { type = "double"; v << *p; }
            else if (auto p = std::any_cast<long double>(&a)) { type = "long double"; v << (double)*p; }
            else if (auto p = std::any_cast<std::string>(&a)) { type = "string"; v << *p; }
            else if (auto p = std::any_cast<std::vector<int>>(&a)) { type = "vector(int)"; for (size_t q = 0; q < p->size(); ++q) v << (q ? " " : "") << (*p)[q]; }
            else if (auto p = std::any_cast<std::vector<double>>(&a)) { type = "vector(double)"; for (size_t q = 0; q < p->size(); ++q) v << (q ? " " : "") << (*p)[q]; }
            else if (auto p = std::any_cast<std::list<std::string>>(&a)) { type = "list(string)"; bool first = true; for (auto &t : *p) { v << (first ? "" : ",") << t; first = false; } }
            out.push_back(prefix + kv.first + "\t" + type + "\t" + v.str());
        }
        for (auto &kv : Sublists_) kv.second->Dump(out, prefix + kv.first + "/");
        if (prefix.empty()) std::sort(out.begin(), out.end());
    }
    void Print(std::ostream &os, unsigned indent = 0) const noexcept
    {
        std::string pad(indent, ' ');
        os << pad << "[" << Name_ << "]\n";
        for (auto &kv : Params_) os << pad << "  " << kv.first << " (" << kv.second.type().name() << ")\n";
        for (auto &kv : Sublists_) kv.second->Print(os, indent + 2);
    }

private:
    template <typename T>
    T &Cast(std::any &a, const key_type &name)
    {
        T *p = std::any_cast<T>(&a);
        PARELAG_TEST_FOR_EXCEPTION(!p, bad_var_cast, "ParameterList: parameter \"" << name << "\" does not hold the requested type.");
        return *p;
    }
    std::string Name_ = "";
    std::unordered_map<key_type, std::any> Params_;
    std::unordered_map<key_type, std::unique_ptr<ParameterList>> Sublists_;
};

/// <ParameterList name=".."> <Parameter name=".." type=".." value=".."/> ... </ParameterList>
/// types: bool,int,double,string,vector(int)/vector_int,vector(double),list(string);
/// unknown parameters are kept (factories ignore what they do not read).
class SimpleXMLParameterListReader
{
public:
    std::unique_ptr<ParameterList> GetParameterList(std::istream &is)
    {
        std::stringstream ss; ss << is.rdbuf();
        return Parse(ss.str());
    }
    std::unique_ptr<ParameterList> Parse(const std::string &text)
    {
        size_t pos = 0;
        std::vector<ParameterList *> stack;
        std::unique_ptr<ParameterList> root;
        while (true)
        {
            size_t lt = text.find('<', pos);
            if (lt == std::string::npos) break;
            if (text.compare(lt, 4, "<!--") == 0) { pos = text.find("-->", lt); PARELAG_ASSERT(pos != std::string::npos); pos += 3; continue; }
            size_t gt = text.find('>', lt);
            PARELAG_TEST_FOR_EXCEPTION(gt == std::string::npos, std::runtime_error, "XML: unterminated tag");
            std::string tag = text.substr(lt + 1, gt - lt - 1);
            pos = gt + 1;
            if (tag.empty() || tag[0] == '?') continue;
            if (tag[0] == '/') { PARELAG_ASSERT(!stack.empty()); stack.pop_back(); continue; }
            bool selfclose = tag.back() == '/';
            if (selfclose) tag.pop_back();
            std::string kind = tag.substr(0, tag.find_first_of(" \t\n"));
            auto attr = Attributes(tag);
            if (kind == "ParameterList")
            {
                ParameterList *pl;
                if (stack.empty()) { root = make_unique<ParameterList>(attr["name"]); pl = root.get(); }
                else pl = &stack.back()->Sublist(attr["name"]);
                if (!selfclose) stack.push_back(pl);
            }
            else if (kind == "Parameter")
            {
                PARELAG_ASSERT(!stack.empty());
                SetTyped(*stack.back(), attr["name"], attr["type"], attr["value"]);
            }
        }
        PARELAG_TEST_FOR_EXCEPTION(!root, std::runtime_error, "XML: no <ParameterList> found");
        return root;
    }

private:
    /// key = "value" pairs of a tag; blanks and line breaks may surround the '=' (examples/example_parameterlists/
    /// 1form_example_parameters.xml writes name = "Type"), values are double-quoted and may contain any character but '"'
    static std::map<std::string, std::string> Attributes(const std::string &tag)
    {
        std::map<std::string, std::string> out;
        const size_t n = tag.size();
        size_t p = tag.find_first_of(" \t\r\n");          // skip the element name
        while (p != std::string::npos && p < n)
        {
            while (p < n && std::isspace((unsigned char)tag[p])) ++p;
            size_t k0 = p;
            while (p < n && !std::isspace((unsigned char)tag[p]) && tag[p] != '=') ++p;
            const std::string key = tag.substr(k0, p - k0);
            while (p < n && std::isspace((unsigned char)tag[p])) ++p;
            if (key.empty() || p >= n || tag[p] != '=') break;
            ++p;
            while (p < n && std::isspace((unsigned char)tag[p])) ++p;
            PARELAG_TEST_FOR_EXCEPTION(p >= n || tag[p] != '"', std::runtime_error, "XML: attribute \"" << key << "\" has no quoted value");
            const size_t q1 = tag.find('"', p + 1);
            PARELAG_TEST_FOR_EXCEPTION(q1 == std::string::npos, std::runtime_error, "XML: unterminated attribute value");
            out[key] = tag.substr(p + 1, q1 - p - 1);
            p = q1 + 1;
        }
        return out;
    }
    static std::vector<std::string> Split(const std::string &s)
    {
        std::vector<std::string> out; std::string cur;
        for (char c : s) { if (c == ',' || c == ' ') { if (!cur.empty()) out.push_back(cur); cur.clear(); } else cur += c; }
        if (!cur.empty()) out.push_back(cur);
        return out;
    }
    static void SetTyped(ParameterList &pl, const std::string &name, const std::string &type, const std::string &value)
    {
        // add_parameter_to_list (ParELAG_SimpleXMLParameterListReader.cpp:207-297): names and values must not be empty;
        // a bool is true iff its value is "true" in any letter case; the full list of arithmetic types
        PARELAG_ASSERT(!name.empty());
        PARELAG_ASSERT(!value.empty());
        if (type == "bool")
        {
            std::string up = value;
            std::transform(up.begin(), up.end(), up.begin(), [](unsigned char c) { return (char)std::toupper(c); });
            pl.Set(name, up == "TRUE");
        }
        else if (type == "char") pl.Set(name, (char)std::stoi(value));
        else if (type == "int") pl.Set(name, std::stoi(value));
        else if (type == "long") pl.Set(name, std::stol(value));
        else if (type == "unsigned long") pl.Set(name, std::stoul(value));
        else if (type == "long long") pl.Set(name, std::stoll(value));
        else if (type == "unsigned long long") pl.Set(name, std::stoull(value));
        else if (type == "size_t") pl.Set(name, (size_t)std::stoull(value));
        else if (type == "float") pl.Set(name, std::stof(value));
        else if (type == "double") pl.Set(name, std::stod(value));
        else if (type == "long double") pl.Set(name, std::stold(value));
        else if (type == "string") pl.Set(name, value);
        else if (type == "vector(int)" || type == "vector_int")
        { std::vector<int> v; for (auto &t : Split(value)) v.push_back(std::stoi(t)); pl.Set(name, v); }
        else if (type == "vector(double)" || type == "vector_double")
        { std::vector<double> v; for (auto &t : Split(value)) v.push_back(std::stod(t)); pl.Set(name, v); }
        else if (type == "list(string)")
        {
            // comma separated, every entry trimmed (trim_string, :363-380)
            std::list<std::string> v; std::string cur;
            auto push = [&v](const std::string &t)
            {
                size_t a = 0, b = t.size();
                while (a < b && std::isspace((unsigned char)t[a])) ++a;
                while (b > a && std::isspace((unsigned char)t[b - 1])) --b;
                v.push_back(t.substr(a, b - a));
            };
            for (char c : value) { if (c == ',') { push(cur); cur.clear(); } else cur += c; }
            if (!cur.empty()) push(cur);
            pl.Set(name, v);
        }
        else PARELAG_TEST_FOR_EXCEPTION(true, std::runtime_error, "XML: unknown parameter type \"" << type << "\"");
    }
};

// ---------------------------------------------------------------- Level
class Level
{
public:
    explicit Level(int id = -1) : ID_(id) {}
    int GetLevelID() const noexcept { return ID_; }
    void SetLevelID(int id) noexcept { ID_ = id; }
    template <typename T> void Set(const std::string &key, T val) { Data_[key] = std::any(std::move(val)); }
    template <typename T> T &Get(const std::string &key)
    {
        auto it = Data_.find(key);
        PARELAG_TEST_FOR_EXCEPTION(it == Data_.end(), std::out_of_range, "Level::Get(): key \"" << key << "\" not found.");
        T *p = std::any_cast<T>(&it->second);
        PARELAG_TEST_FOR_EXCEPTION(!p, bad_var_cast, "Level::Get(): wrong type for key \"" << key << "\".");
        return *p;
    }
    /// the next finer level (ParELAG_Level.hpp: GetPreviousLevel / SetPreviousLevel)
    std::shared_ptr<Level> GetPreviousLevel() { return PreviousLevel_.lock(); }
    void SetPreviousLevel(const std::shared_ptr<Level> &PreviousLevel) { PreviousLevel_ = PreviousLevel; }
    /// overwrite whatever the key holds (Set in the reference refuses to)
    template <typename T> void Reset(const std::string &key, const T &value) { Data_[key] = std::any(value); }
    bool IsKey(const std::string &key) const noexcept { return Data_.count(key) > 0; }
    /// key exists and the stored shared_ptr<mfem::Operator> is non-null
    bool IsValidKey(const std::string &key) const noexcept
    {
        auto it = Data_.find(key);
        if (it == Data_.end()) return false;
        auto p = std::any_cast<std::shared_ptr<mfem::Operator>>(&it->second);
        return p ? (bool)*p : it->second.has_value();
    }
private:
    int ID_;
    std::unordered_map<std::string, std::any> Data_;
    std::weak_ptr<Level> PreviousLevel_;
};

// ---------------------------------------------------------------- Solver
class Solver : public mfem::Solver
{
public:
    Solver(int h, int w, bool iter_mode) : mfem::Solver(h, w, iter_mode) {}
    /// by design: use SetOperator(shared_ptr) (ParELAG_Solver.hpp:57-65,94-97)
    void SetOperator(const mfem::Operator &) final { throw not_implemented_error("Solver::SetOperator(const Operator&): use the shared_ptr overload"); }
    void SetOperator(const std::shared_ptr<mfem::Operator> &op) { _do_set_operator(op); }
    bool IsPreconditioner() const noexcept { return !this->iterative_mode; }
    /// true when Mult() only enqueues device work (no host synchronisation, no allocation after the
    /// first call), i.e. when it may be recorded into a CUDA graph by an enclosing Hierarchy
    virtual bool CaptureSafe() const { return false; }
private:
    virtual void _do_set_operator(const std::shared_ptr<mfem::Operator> &op) = 0;
};

// ---------------------------------------------------------------- SolverState
class DeRhamSequence;
class SolverState
{
public:
    virtual ~SolverState() = default;
    void SetOperator(const std::string &n, const std::shared_ptr<mfem::Operator> &op) { Operators_[n] = op; }
    std::shared_ptr<mfem::Operator> GetOperator(const std::string &n) const noexcept
    { auto it = Operators_.find(n); return it == Operators_.end() ? nullptr : it->second; }
    bool IsOperator(const std::string &n) const noexcept { return Operators_.count(n) > 0; }
    void SetVector(const std::string &n, const std::shared_ptr<mfem::Vector> &v) { Vectors_[n] = v; }
    bool IsVector(const std::string &n) const noexcept { return Vectors_.count(n) > 0; }
    std::shared_ptr<mfem::Vector> GetVector(const std::string &n) const noexcept
    { auto it = Vectors_.find(n); return it == Vectors_.end() ? nullptr : it->second; }
    void SetBoundaryLabels(std::vector<std::vector<int>> labels) noexcept { BoundaryLabels_ = std::move(labels); }
    void SetBoundaryLabels(const std::vector<mfem::Array<int>> &labels)
    {
        BoundaryLabels_.resize(labels.size());
        for (size_t i = 0; i < labels.size(); ++i) BoundaryLabels_[i].assign(labels[i].begin(), labels[i].end());
    }
    std::vector<std::vector<int>> &GetBoundaryLabels() noexcept { return BoundaryLabels_; }
    std::vector<int> &GetBoundaryLabels(int blockID)
    {
        PARELAG_TEST_FOR_EXCEPTION(blockID < 0 || blockID >= (int)BoundaryLabels_.size(), std::out_of_range,
                                   "SolverState::GetBoundaryLabels(): bad block id " << blockID);
        return BoundaryLabels_[blockID];
    }
    void SetDeRhamSequence(const std::shared_ptr<DeRhamSequence> &seq) noexcept { Sequence_ = seq; }
    bool HasDeRhamSequence() const noexcept { return (bool)Sequence_; }
    std::shared_ptr<DeRhamSequence> GetDeRhamSequencePtr() const noexcept { return Sequence_; }
    DeRhamSequence &GetDeRhamSequence() const { PARELAG_ASSERT(Sequence_); return *Sequence_; }
    void SetForms(std::vector<int> forms) noexcept { Forms_ = std::move(forms); }
    std::vector<int> &GetForms() noexcept { return Forms_; }
    void SetExtraParameter(const std::string &n, double v) { Extra_[n] = v; }
    double GetExtraParameter(const std::string &n, double dflt) const noexcept
    { auto it = Extra_.find(n); return it == Extra_.end() ? dflt : it->second; }
    /// values already set in *this win; missing ones are taken from rhs
    virtual void MergeState(const SolverState &rhs)
    {
        for (auto &kv : rhs.Operators_) Operators_.insert(kv);
        for (auto &kv : rhs.Vectors_) Vectors_.insert(kv);
        for (auto &kv : rhs.Extra_) Extra_.insert(kv);
        if (BoundaryLabels_.empty()) BoundaryLabels_ = rhs.BoundaryLabels_;
        if (Forms_.empty()) Forms_ = rhs.Forms_;
        if (!Sequence_) Sequence_ = rhs.Sequence_;
    }
protected:
    std::unordered_map<std::string, std::shared_ptr<mfem::Operator>> Operators_;
    std::unordered_map<std::string, std::shared_ptr<mfem::Vector>> Vectors_;
    std::unordered_map<std::string, double> Extra_;
    std::vector<std::vector<int>> BoundaryLabels_;
    std::vector<int> Forms_;
    std::shared_ptr<DeRhamSequence> Sequence_;
};

class NestedSolverState : public SolverState
{
public:
    void SetSubState(const std::string &n, std::shared_ptr<SolverState> s) { Sub_[n] = std::move(s); }
    std::shared_ptr<SolverState> GetSubState(const std::string &n) noexcept
    { auto it = Sub_.find(n); return it == Sub_.end() ? nullptr : it->second; }
    bool IsSubState(const std::string &n) const noexcept { return Sub_.count(n) > 0; }
    void MergeState(const SolverState &rhs) override
    {
        SolverState::MergeState(rhs);
        if (auto n = dynamic_cast<const NestedSolverState *>(&rhs))
            for (auto &kv : n->Sub_) Sub_.insert(kv);
    }
private:
    std::unordered_map<std::string, std::shared_ptr<SolverState>> Sub_;
};

// ---------------------------------------------------------------- SolverFactory
class SolverLibrary;
class SolverFactory
{
public:
    virtual ~SolverFactory() = default;
    std::unique_ptr<mfem::Solver> BuildSolver(const std::shared_ptr<mfem::Operator> &op, SolverState &state) const
    { return _do_build_solver(op, state); }
    std::unique_ptr<SolverState> GetDefaultState() const { return _do_get_default_state(); }
    void Initialize(const ParameterList &params)
    {
        SetParameters(params);
        SetDefaultParameters();
        _do_initialize(params);
    }
    void SetDefaultParameters() { _do_set_default_parameters(); }
    void SetParameters(const ParameterList &params) { Params_.Merge(params); }
    ParameterList &GetParameters() const { return Params_; }   // Get(name,default) mutates: see HypreSmootherFactory.cpp:33-34
    const SolverLibrary &GetSolverLibrary() const { PARELAG_ASSERT((bool)Lib_); return *Lib_; }
    void SetSolverLibrary(std::shared_ptr<const SolverLibrary> lib) noexcept { Lib_ = std::move(lib); }
    bool HasValidSolverLibrary() const noexcept { return (bool)Lib_; }
private:
    virtual std::unique_ptr<mfem::Solver> _do_build_solver(const std::shared_ptr<mfem::Operator> &op, SolverState &state) const = 0;
    virtual std::unique_ptr<SolverState> _do_get_default_state() const { return make_unique<NestedSolverState>(); }
    virtual void _do_initialize(const ParameterList &params) = 0;
    virtual void _do_set_default_parameters() = 0;
    mutable ParameterList Params_;
    std::shared_ptr<const SolverLibrary> Lib_;
};

// ---------------------------------------------------------------- SolverLibrary
/// name -> {"Type": factory id, "Solver Parameters": {...}} ; factories are created
/// lazily and cached (ParELAG_SolverLibrary.hpp:140-145, .cpp:36-67)
class SolverLibrary : public std::enable_shared_from_this<SolverLibrary>
{
public:
    using creator_type = std::function<std::shared_ptr<SolverFactory>()>;
    static std::shared_ptr<SolverLibrary> CreateLibrary() { auto l = std::shared_ptr<SolverLibrary>(new SolverLibrary()); l->RegisterBuiltins(); return l; }
    static std::shared_ptr<SolverLibrary> CreateLibrary(const ParameterList &pl)
    {
        auto l = CreateLibrary();
        l->Initialize(pl);
        return l;
    }
    void Initialize(const ParameterList &pl)
    {
        for (auto &name : pl.SublistNames())
        {
            const ParameterList &entry = pl.Sublist(name);
            PARELAG_TEST_FOR_EXCEPTION(!entry.IsParameter("Type"), std::runtime_error,
                                       "SolverLibrary: entry \"" << name << "\" has no \"Type\".");
            Entries_[name] = std::make_pair(entry.Get<std::string>("Type"),
                                            entry.IsSublist("Solver Parameters") ? entry.Sublist("Solver Parameters") : ParameterList("Solver Parameters"));
        }
    }
    void AddSolver(const std::string &name, const std::string &type, const ParameterList &params)
    { Entries_[name] = std::make_pair(type, params); Cache_.erase(name); }
    void AddSolver(const std::string &name, std::shared_ptr<SolverFactory> fact) { Cache_[name] = std::move(fact); }
    bool IsSolver(const std::string &name) const noexcept { return Entries_.count(name) > 0 || Cache_.count(name) > 0; }
    void RegisterFactoryType(const std::string &type, creator_type c) { Creators_[type] = std::move(c); }
    /// the reference's registry interface (ParELAG_SolverLibrary.hpp:177-224): names of the solvers in the library, and
    /// adding / removing / listing factory types (a user type: AddNewSolverFactory("SuperCool", [] { return
    /// std::make_shared<SuperCoolSolverFactory>(); }))
    std::list<std::string> GetSolverNames() const
    {
        std::list<std::string> ret;
        for (auto &e : Entries_) ret.push_back(e.first);
        for (auto &c : Cache_) if (!Entries_.count(c.first)) ret.push_back(c.first);
        return ret;
    }
    bool AddNewSolverFactory(const std::string &id, creator_type builder) { return Creators_.emplace(id, std::move(builder)).second; }
    bool RemoveSolverFactory(const std::string &id) { return Creators_.erase(id) > 0; }
    std::list<std::string> GetSolverFactoryNames() const
    {
        std::list<std::string> ret;
        for (auto &c : Creators_) ret.push_back(c.first);
        return ret;
    }
    std::shared_ptr<SolverFactory> GetSolverFactory(const std::string &name) const
    {
        auto c = Cache_.find(name);
        if (c != Cache_.end()) return c->second;
        auto e = Entries_.find(name);
        PARELAG_TEST_FOR_EXCEPTION(e == Entries_.end(), std::out_of_range,
                                   "SolverLibrary::GetSolverFactory(): solver \"" << name << "\" is not in the library.");
        auto cr = Creators_.find(e->second.first);
        PARELAG_TEST_FOR_EXCEPTION(cr == Creators_.end(), std::runtime_error,
                                   "SolverLibrary: unknown factory type \"" << e->second.first << "\" for solver \"" << name << "\".");
        auto fact = cr->second();
        fact->SetSolverLibrary(shared_from_this());
        Cache_[name] = fact;          // before Initialize: nested lookups may recurse
        try { fact->Initialize(e->second.second); }
        catch (...) { Cache_.erase(name); throw; }      // a factory whose initialisation failed must not be handed out later
        return fact;
    }
private:
    SolverLibrary() = default;
    void RegisterBuiltins();          // defined in parelag_solvers.hpp
    std::unordered_map<std::string, std::pair<std::string, ParameterList>> Entries_;
    mutable std::unordered_map<std::string, std::shared_ptr<SolverFactory>> Cache_;
    std::unordered_map<std::string, creator_type> Creators_;
};

/// replaces parelag::mpi_session (src/utilities/mpiUtils.cpp:23-31): one rank <-> one GPU
class mpi_session
{
public:
    mpi_session(int rank = 0, int nranks = 1, int device = 0, const void *nccl_id = nullptr)
    {
        PARELAG_TEST_FOR_EXCEPTION(Device::Ctx() != nullptr, std::runtime_error, "mpi_session: a session already exists");
        PE_CALL(pe_ctx_create(rank, nranks, device, nccl_id, &Device::Ctx()));
    }
    ~mpi_session() { pe_ctx_destroy(Device::Ctx()); Device::Ctx() = nullptr; }
};
} // namespace parelag
