// gtsam_facade.h -- header-only, GTSAM-shaped front end of the B200 batch optimiser (include/asc.h).
//
// Purpose: let ArmSlamCalib::OptimizeStep (src/ArmSlamCalib.cpp:429-533) keep its source for the two batch modes:
//     LevenbergMarquardtParams params; params.linearSolverType = LevenbergMarquardtParams::SEQUENTIAL_CHOLESKY;
//     LevenbergMarquardtOptimizer optimizer(*graph, currentEstimate, params);          // :460-462
//     currentEstimate = optimizer.optimize();  optimizer.error();                      // :463-465
//     DoglegParams doglegParams; DoglegOptimizer optimizer(*graph, currentEstimate, doglegParams);   // :442-447
//     catch (gtsam::IndeterminantLinearSystemException& e) { keyFormatter(e.nearbyVariable()) }     // :451-455
// and the graph/value funnels AddFactor (:1916-1929) / AddValue (:1931-1941) with the factor types the repo owns.
//
// What it is NOT: a GTSAM re-implementation.  It provides exactly the symbols that path touches, holds
// factors as plain data, and packs them -- by dynamic_cast, the way the reference inspects its own graph
// (ArmSlamCalib.cpp:1893, :2066, :2167) -- into the flat arrays of asc_set_problem.  All numerics run on the GPU.
//
// The DART skeleton handle the reference's factors carry (RobotProjectionFactor.h:33-35) is replaced by
// asc_facade::SerialChain; INTEGRATION.md shows the 15 lines that fill it from a dart::dynamics::MetaSkeleton.
// Without Eigen in this container gtsam::Vector is a small std::vector wrapper; with Eigen present a maintainer
// would typedef it to Eigen::VectorXd (only operator(), size() and data() are used here).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../asc.h"

namespace asc_facade {

// Serial revolute chain description = what asc_set_chain needs (DART convention)
struct SerialChain {
    std::vector<double> t_p2j, axis, t_c2j;   // [N][12], [N][3], [N][12]
    std::vector<uint8_t> continuous;          // RobotConfig::IsContinuous, RobotConfig.h:72-77
    int dof() const { return (int)(axis.size() / 3); }
};
typedef std::shared_ptr<SerialChain> SerialChainPtr;

}  // namespace asc_facade

namespace gtsam {

typedef std::uint64_t Key;

class Vector {
public:
    Vector() {}
    explicit Vector(size_t n, double v = 0.0) : d_(n, v) {}
    Vector(std::initializer_list<double> l) : d_(l) {}
    static Vector Constant(size_t n, double v) { return Vector(n, v); }
    double& operator()(size_t i) { return d_[i]; }
    double operator()(size_t i) const { return d_[i]; }
    size_t size() const { return d_.size(); }
    size_t rows() const { return d_.size(); }
    const double* data() const { return d_.data(); }
    double* data() { return d_.data(); }
private:
    std::vector<double> d_;
};

// dense row-major matrix; only what ArmSlamCalib.cpp does with a marginal covariance: rows(), operator()(i, j) (:976-985)
class Matrix {
public:
    Matrix() : r_(0), c_(0) {}
    Matrix(size_t r, size_t c, double v = 0.0) : r_(r), c_(c), d_(r * c, v) {}
    double& operator()(size_t i, size_t j) { return d_[i * c_ + j]; }
    double operator()(size_t i, size_t j) const { return d_[i * c_ + j]; }
    size_t rows() const { return r_; }
    size_t cols() const { return c_; }
    const double* data() const { return d_.data(); }
    double* data() { return d_.data(); }
private:
    size_t r_, c_;
    std::vector<double> d_;
};

// Symbol(char, index), ArmSlamCalib.h:284-286: 'K' extrinsic, 'q' configuration, 'l' landmark
class Symbol {
public:
    Symbol(unsigned char c, std::uint64_t j) : c_(c), j_(j) {}
    explicit Symbol(Key k) : c_((unsigned char)(k >> 56)), j_(k & 0x00ffffffffffffffull) {}
    operator Key() const { return key(); }
    Key key() const { return ((Key)c_ << 56) | j_; }
    unsigned char chr() const { return c_; }
    std::uint64_t index() const { return j_; }
private:
    unsigned char c_;
    std::uint64_t j_;
};
inline std::string DefaultKeyFormatter(Key k) { Symbol s(k); std::ostringstream o; o << s.chr() << s.index(); return o.str(); }
typedef std::string (*KeyFormatter)(Key);

struct Point2 { double x_, y_; Point2(double x = 0, double y = 0) : x_(x), y_(y) {} double x() const { return x_; } double y() const { return y_; } };
struct Point3 { double x_, y_, z_; Point3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {}
                double x() const { return x_; } double y() const { return y_; } double z() const { return z_; }
                Point3 operator-(const Point3& o) const { return Point3(x_ - o.x_, y_ - o.y_, z_ - o.z_); }     // ComputeError, ArmSlamCalib.cpp:1377
                Point3 operator+(const Point3& o) const { return Point3(x_ + o.x_, y_ + o.y_, z_ + o.z_); }
                double norm() const { return std::sqrt(x_ * x_ + y_ * y_ + z_ * z_); } };
struct Quaternion { double x_, y_, z_, w_; double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double w() const { return w_; } };

class Rot3 {
public:
    Rot3() { std::memset(R, 0, sizeof R); R[0] = R[4] = R[8] = 1; }
    static Rot3 identity() { return Rot3(); }
    static Rot3 rodriguez(double wx, double wy, double wz);   // sim_genrobot.cpp:67
    Quaternion toQuaternion() const {                          // PrintSimErrors, ArmSlamCalib.cpp:2224 (Eigen's matrix -> quaternion branches)
        Quaternion q;
        const double t = R[0] + R[4] + R[8];
        if (t > 0) {
            const double s = std::sqrt(t + 1.0) * 2;
            q.w_ = 0.25 * s; q.x_ = (R[7] - R[5]) / s; q.y_ = (R[2] - R[6]) / s; q.z_ = (R[3] - R[1]) / s;
        } else {
            int i = 0;
            if (R[4] > R[0]) i = 1;
            if (R[8] > R[4 * i]) i = 2;
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            const double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0) * 2;
            double v[3];
            v[i] = 0.25 * s; v[j] = (R[3 * j + i] + R[3 * i + j]) / s; v[k] = (R[3 * k + i] + R[3 * i + k]) / s;
            q.x_ = v[0]; q.y_ = v[1]; q.z_ = v[2]; q.w_ = (R[3 * k + j] - R[3 * j + k]) / s;
        }
        return q;
    }
    double R[9];   // row-major
};

class Pose3 {
public:
    Pose3() {}
    Pose3(const Rot3& r, const Point3& t) : R_(r), t_(t) {}
    static Pose3 identity() { return Pose3(); }
    const Rot3& rotation() const { return R_; }
    const Point3& translation() const { return t_; }
    void to12(double* m) const {
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m[4 * r + c] = R_.R[3 * r + c]; }
        m[3] = t_.x(); m[7] = t_.y(); m[11] = t_.z();
    }
    static Pose3 from12(const double* m) {
        Rot3 r; for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) r.R[3 * i + c] = m[4 * i + c];
        return Pose3(r, Point3(m[3], m[7], m[11]));
    }
private:
    Rot3 R_;
    Point3 t_;
};

inline Rot3 Rot3::rodriguez(double wx, double wy, double wz) {
    Rot3 o;
    const double th = std::sqrt(wx * wx + wy * wy + wz * wz);
    if (th < 1e-10) return o;
    const double x = wx / th, y = wy / th, z = wz / th, s = std::sin(th), c = std::cos(th), v = 1 - c;
    const double R[9] = {c + x * x * v, x * y * v - z * s, x * z * v + y * s, y * x * v + z * s, c + y * y * v, y * z * v - x * s,
                         z * x * v - y * s, z * y * v + x * s, c + z * z * v};
    std::memcpy(o.R, R, sizeof R);
    return o;
}

class Cal3_S2 {
public:
    Cal3_S2(double fx, double fy, double s, double u0, double v0) : fx_(fx), fy_(fy), s_(s), u0_(u0), v0_(v0) {}
    double fx() const { return fx_; } double fy() const { return fy_; } double skew() const { return s_; }
    double px() const { return u0_; } double py() const { return v0_; }
private:
    double fx_, fy_, s_, u0_, v0_;
};

// RobotConfig value (include/arm_slam_calib/RobotConfig.h): the joint vector of one keyframe
class RobotConfig {
public:
    RobotConfig() {}
    RobotConfig(const Vector& q, const asc_facade::SerialChainPtr& arm) : q_(q), arm_(arm) {}
    const Vector& getQ() const { return q_; }
    size_t dim() const { return q_.size(); }
    const asc_facade::SerialChainPtr& arm() const { return arm_; }
private:
    Vector q_;
    asc_facade::SerialChainPtr arm_;
};

// ---- Values: key-ordered heterogeneous container (insert / erase / find / end / begin / at<T> / exists)
class Values {
public:
    struct ValueBase { virtual ~ValueBase() {} };
    template <class T> struct Holder : ValueBase { T v; explicit Holder(const T& x) : v(x) {} };
    struct KeyValuePair { Key key; const ValueBase& value; };
    typedef std::map<Key, std::shared_ptr<ValueBase> > Map;
    class const_iterator {
    public:
        explicit const_iterator(Map::const_iterator it) : it_(it) {}
        bool operator!=(const const_iterator& o) const { return it_ != o.it_; }
        bool operator==(const const_iterator& o) const { return it_ == o.it_; }
        const_iterator& operator++() { ++it_; return *this; }
        const_iterator operator++(int) { const_iterator t = *this; ++it_; return t; }
        struct Proxy { Key key; const ValueBase& value; const Proxy* operator->() const { return this; } };
        Proxy operator*() const { return Proxy{it_->first, *it_->second}; }
        Proxy operator->() const { return Proxy{it_->first, *it_->second}; }
    private:
        Map::const_iterator it_;
    };
    template <class T> void insert(Key k, const T& v) {
        if (!m_.insert(std::make_pair(k, std::shared_ptr<ValueBase>(new Holder<T>(v)))).second) throw std::invalid_argument("ValuesKeyAlreadyExists " + DefaultKeyFormatter(k));
    }
    template <class T> void update(Key k, const T& v) { m_[k] = std::shared_ptr<ValueBase>(new Holder<T>(v)); }
    void erase(Key k) { m_.erase(k); }
    bool exists(Key k) const { return m_.count(k) != 0; }
    const_iterator find(Key k) const { return const_iterator(m_.find(k)); }
    const_iterator begin() const { return const_iterator(m_.begin()); }
    const_iterator end() const { return const_iterator(m_.end()); }
    size_t size() const { return m_.size(); }
    template <class T> const T& at(Key k) const {
        Map::const_iterator it = m_.find(k);
        if (it == m_.end()) throw std::out_of_range("ValuesKeyDoesNotExist " + DefaultKeyFormatter(k));
        const Holder<T>* h = dynamic_cast<const Holder<T>*>(it->second.get());
        if (!h) throw std::invalid_argument("ValuesIncorrectType " + DefaultKeyFormatter(k));
        return h->v;
    }
    const Map& map() const { return m_; }
private:
    Map m_;
};

// ---- noise models actually used: Diagonal::Sigmas, Robust::Create(Cauchy, Diagonal)  (ArmSlamCalib.cpp:672-705)
namespace noiseModel {
class Base { public: virtual ~Base() {} };
typedef std::shared_ptr<Base> shared_ptr;
class Diagonal : public Base {
public:
    typedef std::shared_ptr<Diagonal> shared_ptr;
    static shared_ptr Sigmas(const Vector& s) { shared_ptr p(new Diagonal()); p->sigmas_ = s; return p; }
    const Vector& sigmas() const { return sigmas_; }
private:
    Vector sigmas_;
};
namespace mEstimator {
class Cauchy {
public:
    typedef std::shared_ptr<Cauchy> shared_ptr;
    explicit Cauchy(double k) : k_(k) {}
    static shared_ptr Create(double k) { return shared_ptr(new Cauchy(k)); }
    double k() const { return k_; }
private:
    double k_;
};
}  // namespace mEstimator
class Robust : public Base {
public:
    typedef std::shared_ptr<Robust> shared_ptr;
    static shared_ptr Create(const mEstimator::Cauchy::shared_ptr& r, const Diagonal::shared_ptr& n) { shared_ptr p(new Robust()); p->robust_ = r; p->noise_ = n; return p; }
    const mEstimator::Cauchy::shared_ptr& robust() const { return robust_; }
    const Diagonal::shared_ptr& noise() const { return noise_; }
private:
    mEstimator::Cauchy::shared_ptr robust_;
    Diagonal::shared_ptr noise_;
};
}  // namespace noiseModel
typedef noiseModel::shared_ptr SharedNoiseModel;

// ---- factors (plain data; evaluation happens in the CUDA kernels)
class NonlinearFactor {
public:
    typedef std::shared_ptr<NonlinearFactor> shared_ptr;
    virtual ~NonlinearFactor() {}
    const std::vector<Key>& keys() const { return keys_; }
    const SharedNoiseModel& noiseModel() const { return noise_; }
protected:
    std::vector<Key> keys_;
    SharedNoiseModel noise_;
};

template <class T> class PriorFactor : public NonlinearFactor {
public:
    PriorFactor(Key key, const T& prior, const SharedNoiseModel& model) : prior_(prior) { keys_.push_back(key); noise_ = model; }
    const T& prior() const { return prior_; }
    Key key() const { return keys_[0]; }
private:
    T prior_;
};

class EncoderFactor : public NonlinearFactor {   // include/arm_slam_calib/EncoderFactor.h
public:
    EncoderFactor(Key robotConfig, const Vector& measurement, const SharedNoiseModel& model, bool hasDeadBand, double deadBand)
        : encoders(measurement), deadBand_(deadBand), hasDeadBand_(hasDeadBand) { keys_.push_back(robotConfig); noise_ = model; }
    const Vector& measurement() const { return encoders; }
    bool hasDeadBand() const { return hasDeadBand_; }
    double deadBand() const { return deadBand_; }
private:
    Vector encoders;
    double deadBand_;
    bool hasDeadBand_;
};

class DriftFactor : public NonlinearFactor {   // include/arm_slam_calib/DriftFactor.h:18-61
public:
    DriftFactor(Key q0, Key q1, const Vector& q0Encoders_, const Vector& q1Encoders_, const SharedNoiseModel& model)
        : q0Encoders(q0Encoders_), q1Encoders(q1Encoders_) { keys_.push_back(q0); keys_.push_back(q1); noise_ = model; }
    const Vector& q0Measurement() const { return q0Encoders; }
    const Vector& q1Measurement() const { return q1Encoders; }
private:
    Vector q0Encoders, q1Encoders;
};

class VelocityFactor : public NonlinearFactor {   // include/arm_slam_calib/VelocityFactor.h:20-35
public:
    VelocityFactor(Key q0, Key q1, const Vector& qDot_, const SharedNoiseModel& model) : qDot(qDot_) { keys_.push_back(q0); keys_.push_back(q1); noise_ = model; }
    const Vector& measurement() const { return qDot; }
private:
    Vector qDot;
};

template <class CameraCalibration> class RobotProjectionFactor : public NonlinearFactor {   // RobotProjectionFactor.h:52-70
public:
    RobotProjectionFactor(const Point2& measured, const SharedNoiseModel& model, Key robotKey, Key landmarkKey, Key extrinsicKey,
                          const asc_facade::SerialChainPtr& robot, const void* /*linkBody*/, const void* /*robotMutex*/,
                          const std::shared_ptr<CameraCalibration>& calibration, size_t traj, size_t landmark,
                          bool /*throwCheirality*/ = false, bool /*verboseCheirality*/ = false)
        : measurement(measured), K(calibration), robot_(robot), trajIndex(traj), landmarkIndex(landmark) {
        keys_.push_back(robotKey); keys_.push_back(landmarkKey); keys_.push_back(extrinsicKey); noise_ = model;
    }
    const Point2 getMeasurement() const { return measurement; }
    size_t getTrajIndex() const { return trajIndex; }
    size_t getLandmarkIndex() const { return landmarkIndex; }
    const std::shared_ptr<CameraCalibration>& calibration() const { return K; }
    const asc_facade::SerialChainPtr& robot() const { return robot_; }
private:
    Point2 measurement;
    std::shared_ptr<CameraCalibration> K;
    asc_facade::SerialChainPtr robot_;
    size_t trajIndex, landmarkIndex;
};

class NonlinearFactorGraph {
public:
    typedef std::vector<NonlinearFactor::shared_ptr>::const_iterator const_iterator;
    void add(const NonlinearFactor::shared_ptr& f) { factors_.push_back(f); }
    void push_back(const NonlinearFactor::shared_ptr& f) { factors_.push_back(f); }
    const_iterator begin() const { return factors_.begin(); }
    const_iterator end() const { return factors_.end(); }
    size_t size() const { return factors_.size(); }
    void resize(size_t n) { factors_.resize(n); }
    const NonlinearFactor::shared_ptr& at(size_t i) const { return factors_.at(i); }
    // `*newGraph = graph->clone();` (ArmSlamCalib.cpp:2209).  The facade's factors are immutable data holders, so sharing them is a deep
    // copy in effect.
    NonlinearFactorGraph clone() const { return *this; }
    // graphviz dot of the factor graph, `graph->saveGraph(os, currentEstimate)` (ArmSlamCalib.cpp:1910-1913): one node per variable,
    // one point per factor, edges factor -- variable
    void saveGraph(std::ostream& os, const Values& values, KeyFormatter fmt = DefaultKeyFormatter) const {
        os << "graph {\n  size=\"5,5\";\n\n";
        for (Values::const_iterator it = values.begin(); it != values.end(); ++it)
            os << "  var" << (*it).key << "[label=\"" << fmt((*it).key) << "\"];\n";
        os << "\n";
        for (size_t i = 0; i < factors_.size(); ++i) {
            if (!factors_[i]) continue;
            os << "  factor" << i << "[label=\"\", shape=point];\n";
            for (size_t k = 0; k < factors_[i]->keys().size(); ++k) os << "  var" << factors_[i]->keys()[k] << "--factor" << i << ";\n";
        }
        os << "}\n";
    }
private:
    std::vector<NonlinearFactor::shared_ptr> factors_;
};

class IndeterminantLinearSystemException : public std::exception {
public:
    explicit IndeterminantLinearSystemException(Key j) : j_(j) {}
    Key nearbyVariable() const { return j_; }
    const char* what() const noexcept override { return "Indeterminant linear system detected while working near a variable"; }
private:
    Key j_;
};

struct NonlinearOptimizerParams {
    size_t maxIterations = 100;
    double relativeErrorTol = 1e-5, absoluteErrorTol = 1e-5, errorTol = 0.0;
    enum LinearSolverType { MULTIFRONTAL_CHOLESKY, MULTIFRONTAL_QR, SEQUENTIAL_CHOLESKY, SEQUENTIAL_QR, Iterative, CHOLMOD };
    LinearSolverType linearSolverType = MULTIFRONTAL_CHOLESKY;   // accepted and ignored: the GPU path has one solver
};
struct LevenbergMarquardtParams : NonlinearOptimizerParams {
    double lambdaInitial = 1e-5, lambdaFactor = 10.0, lambdaUpperBound = 1e5, lambdaLowerBound = 0.0, minModelFidelity = 1e-3;
};
struct DoglegParams : NonlinearOptimizerParams { double deltaInitial = 1.0; };

namespace facade_detail {

inline void check(int rc, asc_handle* h) {
    if (rc == ASC_OK) return;
    if (rc == ASC_ERR_INDETERMINATE) {
        char kind = '?'; int64_t idx = 0;
        asc_indeterminate_key(h, &kind, &idx);
        throw IndeterminantLinearSystemException(Symbol((unsigned char)kind, (std::uint64_t)idx));
    }
    throw std::runtime_error(std::string("arm_slam_calib_b200: ") + asc_last_error());
}

// Flattens (graph, values) into the C ABI and runs one optimize().  Owns the device handle.
class BatchOptimizer {
public:
    BatchOptimizer(const NonlinearFactorGraph& graph, const Values& initial, int mode, const asc_params& prm)
        : values_(initial), mode_(mode), h_(nullptr), error_(0.0), iterations_(0) {
        pack(graph, initial, prm);
    }
    ~BatchOptimizer() { if (h_) asc_destroy(h_); }
    BatchOptimizer(const BatchOptimizer&) = delete;
    BatchOptimizer& operator=(const BatchOptimizer&) = delete;

    const Values& optimize() {
        std::vector<asc_iter_log> log(std::max(1, (int)prm_.max_iterations));
        int32_t n = 0;
        const int rc = asc_optimize(h_, mode_, log.data(), (int32_t)log.size(), &n, &error_);
        iterations_ = n;
        check(rc, h_);   // on ASC_ERR_INDETERMINATE the estimate is left unchanged, like the reference (:451-455)
        std::vector<double> q(pose_keys_.size() * N_), l(lm_keys_.size() * 3), x(12);
        check(asc_get_values(h_, q.data(), l.data(), x.data()), h_);
        for (size_t p = 0; p < pose_keys_.size(); ++p) {
            Vector v(N_);
            for (int i = 0; i < N_; ++i) v(i) = q[p * N_ + i];
            values_.update(pose_keys_[p], RobotConfig(v, chain_));
        }
        for (size_t j = 0; j < lm_keys_.size(); ++j) values_.update(lm_keys_[j], Point3(l[3 * j], l[3 * j + 1], l[3 * j + 2]));
        values_.update(extrinsic_key_, Pose3::from12(x.data()));
        return values_;
    }
    double error() const { return error_; }
    int iterations() const { return iterations_; }
    const Values& values() const { return values_; }
    Key extrinsicKey() const { return extrinsic_key_; }
    // K0 block of (A^T A)^-1 at the values the handle currently holds (the initial ones before optimize(), the result after)
    Matrix extrinsicMarginal() {
        Matrix m(6, 6);
        check(asc_extrinsic_marginal(h_, m.data(), nullptr), h_);
        return m;
    }

private:
    void pack(const NonlinearFactorGraph& graph, const Values& initial, const asc_params& prm_in) {
        prm_ = prm_in;
        // ---- variables: 'q' keyframes, 'l' landmarks, 'K' extrinsic, in key order (GTSAM Values order)
        std::map<Key, int> pose_of, lm_of;
        bool have_k = false;
        for (Values::const_iterator it = initial.begin(); it != initial.end(); ++it) {
            const Symbol s((*it).key);
            if (s.chr() == 'q') { pose_of[(*it).key] = (int)pose_keys_.size(); pose_keys_.push_back((*it).key); }
            else if (s.chr() == 'l') { lm_of[(*it).key] = (int)lm_keys_.size(); lm_keys_.push_back((*it).key); }
            else if (s.chr() == 'K') { extrinsic_key_ = (*it).key; have_k = true; }
            else throw std::invalid_argument("unsupported variable " + DefaultKeyFormatter((*it).key));
        }
        if (!have_k || pose_keys_.empty()) throw std::invalid_argument("the graph needs the extrinsic K0 and at least one configuration");
        chain_ = initial.at<RobotConfig>(pose_keys_[0]).arm();
        if (!chain_) throw std::invalid_argument("RobotConfig carries no SerialChain");
        N_ = chain_->dof();
        const size_t P = pose_keys_.size(), L = lm_keys_.size();
        // ---- factors
        std::vector<int32_t> pose_idx, lm_idx;
        std::vector<double> uv, enc(P * N_, 0.0), lm_prior(L * 3, 0.0), x_prior(12, 0.0);
        std::vector<uint8_t> has_enc(P, 0), has_lm_prior(L, 0), has_drift(P, 0);
        std::vector<double> drift_meas(P * N_, 0.0), sigma_drift(N_, 0.05);
        std::vector<uint8_t> has_vel(P, 0);
        std::vector<double> vel_meas(P * N_, 0.0), sigma_vel(N_, 0.01);
        bool have_drift = false, have_vel = false;
        bool have_x_prior = false, have_cam = false;
        double fx = 0, fy = 0, cx = 0, cy = 0;
        for (NonlinearFactorGraph::const_iterator it = graph.begin(); it != graph.end(); ++it) {
            const NonlinearFactor* f = it->get();
            if (!f) continue;
            if (const RobotProjectionFactor<Cal3_S2>* pf = dynamic_cast<const RobotProjectionFactor<Cal3_S2>*>(f)) {
                pose_idx.push_back(pose_of.at(pf->keys()[0]));
                lm_idx.push_back(lm_of.at(pf->keys()[1]));
                uv.push_back(pf->getMeasurement().x()); uv.push_back(pf->getMeasurement().y());
                const noiseModel::Robust* rb = dynamic_cast<const noiseModel::Robust*>(pf->noiseModel().get());
                if (!rb) throw std::invalid_argument("projection factors need Robust(Cauchy, Diagonal) noise (ArmSlamCalib.cpp:704)");
                prm_.sigma_px = rb->noise()->sigmas()(0);
                prm_.cauchy_k = rb->robust()->k();
                if (!have_cam) { fx = pf->calibration()->fx(); fy = pf->calibration()->fy(); cx = pf->calibration()->px(); cy = pf->calibration()->py(); have_cam = true; }
            } else if (const EncoderFactor* ef = dynamic_cast<const EncoderFactor*>(f)) {
                const int p = pose_of.at(ef->keys()[0]);
                has_enc[p] = 1;
                for (int i = 0; i < N_; ++i) enc[(size_t)p * N_ + i] = ef->measurement()(i);
                const noiseModel::Diagonal* d = dynamic_cast<const noiseModel::Diagonal*>(ef->noiseModel().get());
                if (!d) throw std::invalid_argument("encoder factors need Diagonal noise (ArmSlamCalib.cpp:677)");
                for (int i = 0; i < N_; ++i) prm_.sigma_encoder[i] = d->sigmas()(i);
                prm_.use_dead_band = ef->hasDeadBand() ? 1 : 0;
                prm_.dead_band = ef->deadBand();
            } else if (const PriorFactor<Point3>* lp = dynamic_cast<const PriorFactor<Point3>*>(f)) {
                const int j = lm_of.at(lp->key());
                has_lm_prior[j] = 1;
                lm_prior[3 * j] = lp->prior().x(); lm_prior[3 * j + 1] = lp->prior().y(); lm_prior[3 * j + 2] = lp->prior().z();
                const noiseModel::Robust* rb = dynamic_cast<const noiseModel::Robust*>(lp->noiseModel().get());
                if (!rb) throw std::invalid_argument("landmark priors need Robust(Cauchy, Diagonal) noise (ArmSlamCalib.cpp:705)");
                prm_.sigma_landmark = rb->noise()->sigmas()(0);
            } else if (const PriorFactor<Pose3>* xp = dynamic_cast<const PriorFactor<Pose3>*>(f)) {
                xp->prior().to12(x_prior.data());
                const noiseModel::Diagonal* d = dynamic_cast<const noiseModel::Diagonal*>(xp->noiseModel().get());
                if (!d) throw std::invalid_argument("the extrinsic prior needs Diagonal noise (ArmSlamCalib.cpp:684)");
                for (int i = 0; i < 6; ++i) prm_.sigma_extrinsic[i] = d->sigmas()(i);   // tangent order, quirk of :684 preserved
                have_x_prior = true;
            } else if (const DriftFactor* df = dynamic_cast<const DriftFactor*>(f)) {
                // AddDriftFactor (ArmSlamCalib.cpp:1653-1668) links consecutive configurations q_{i-1}, q_i
                const int p0 = pose_of.at(df->keys()[0]), p1 = pose_of.at(df->keys()[1]);
                if (p1 != p0 + 1) throw std::invalid_argument("DriftFactor must link consecutive configurations (ArmSlamCalib.cpp:1668)");
                if (has_drift[p1]) throw std::invalid_argument("duplicate DriftFactor on " + DefaultKeyFormatter(df->keys()[1]));
                has_drift[p1] = 1; have_drift = true;
                for (int i = 0; i < N_; ++i) drift_meas[(size_t)p1 * N_ + i] = df->q1Measurement()(i) - df->q0Measurement()(i);
                const noiseModel::Diagonal* d = dynamic_cast<const noiseModel::Diagonal*>(df->noiseModel().get());
                if (!d) throw std::invalid_argument("drift factors need Diagonal noise (ArmSlamCalib.cpp:678)");
                for (int i = 0; i < N_; ++i) sigma_drift[i] = d->sigmas()(i);
            } else if (const VelocityFactor* vf = dynamic_cast<const VelocityFactor*>(f)) {
                // AddVelocityFactor (ArmSlamCalib.cpp:1671-1686) links consecutive configurations q_{i-1}, q_i with velocities[i-1]
                const int p0 = pose_of.at(vf->keys()[0]), p1 = pose_of.at(vf->keys()[1]);
                if (p1 != p0 + 1) throw std::invalid_argument("VelocityFactor must link consecutive configurations (ArmSlamCalib.cpp:1685)");
                if (has_vel[p1]) throw std::invalid_argument("duplicate VelocityFactor on " + DefaultKeyFormatter(vf->keys()[1]));
                has_vel[p1] = 1; have_vel = true;
                for (int i = 0; i < N_; ++i) vel_meas[(size_t)p1 * N_ + i] = vf->measurement()(i);
                const noiseModel::Diagonal* d = dynamic_cast<const noiseModel::Diagonal*>(vf->noiseModel().get());
                if (!d) throw std::invalid_argument("velocity factors need Diagonal noise (ArmSlamCalib.cpp:679)");
                for (int i = 0; i < N_; ++i) sigma_vel[i] = d->sigmas()(i);
            } else {
                throw std::invalid_argument("factor type outside the batch path");
            }
        }
        if (!have_x_prior) throw std::invalid_argument("the graph has no PriorFactor<Pose3> on K0 (ArmSlamCalib.cpp:685)");
        for (size_t j = 0; j < L; ++j)
            if (!has_lm_prior[j]) throw std::invalid_argument("landmark " + DefaultKeyFormatter(lm_keys_[j]) + " has no PriorFactor<Point3> (ArmSlamCalib.cpp:1648)");
        if (!have_cam) { fx = fy = 1; }
        // ---- values
        std::vector<double> q(P * N_), l(L * 3), x(12);
        for (size_t p = 0; p < P; ++p) { const Vector& v = initial.at<RobotConfig>(pose_keys_[p]).getQ(); for (int i = 0; i < N_; ++i) q[p * N_ + i] = v(i); }
        for (size_t j = 0; j < L; ++j) { const Point3& pt = initial.at<Point3>(lm_keys_[j]); l[3 * j] = pt.x(); l[3 * j + 1] = pt.y(); l[3 * j + 2] = pt.z(); }
        initial.at<Pose3>(extrinsic_key_).to12(x.data());
        // ---- device
        check(asc_create(&prm_, &h_), h_);
        check(asc_set_chain(h_, N_, chain_->t_p2j.data(), chain_->axis.data(), chain_->t_c2j.data(),
                            chain_->continuous.empty() ? nullptr : chain_->continuous.data(), nullptr), h_);
        check(asc_set_camera(h_, fx, fy, cx, cy), h_);
        check(asc_set_problem(h_, (int32_t)P, (int32_t)L, (int64_t)pose_idx.size(), pose_idx.data(), lm_idx.data(), uv.data(), enc.data(),
                              has_enc.data(), lm_prior.data(), x_prior.data()), h_);
        if (have_drift) check(asc_set_drift(h_, has_drift.data(), drift_meas.data(), sigma_drift.data()), h_);
        if (have_vel) check(asc_set_velocity(h_, has_vel.data(), vel_meas.data(), sigma_vel.data()), h_);
        check(asc_set_values(h_, q.data(), l.data(), x.data()), h_);
    }

    Values values_;
    int mode_;
    asc_handle* h_;
    asc_params prm_;
    asc_facade::SerialChainPtr chain_;
    int N_ = 0;
    std::vector<Key> pose_keys_, lm_keys_;
    Key extrinsic_key_ = 0;
    double error_;
    int iterations_;
};

inline asc_params base_params(const NonlinearOptimizerParams& p) {
    asc_params a;
    asc_default_params(&a);
    a.max_iterations = (int32_t)p.maxIterations;
    a.relative_error_tol = p.relativeErrorTol; a.absolute_error_tol = p.absoluteErrorTol; a.error_tol = p.errorTol;
    return a;
}

}  // namespace facade_detail

class LevenbergMarquardtOptimizer {
public:
    LevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues, const LevenbergMarquardtParams& params = LevenbergMarquardtParams())
        : impl_(graph, initialValues, ASC_BATCH_LM, make(params)) {}
    const Values& optimize() { return impl_.optimize(); }
    double error() const { return impl_.error(); }
    int iterations() const { return impl_.iterations(); }
    const Values& values() const { return impl_.values(); }
private:
    static asc_params make(const LevenbergMarquardtParams& p) {
        asc_params a = facade_detail::base_params(p);
        a.lambda_initial = p.lambdaInitial; a.lambda_factor = p.lambdaFactor; a.lambda_upper_bound = p.lambdaUpperBound;
        a.lambda_lower_bound = p.lambdaLowerBound; a.min_model_fidelity = p.minModelFidelity;
        return a;
    }
    facade_detail::BatchOptimizer impl_;
};

// gtsam::Marginals, the batch counterpart of isam2.marginalCovariance(ExtrinsicSymbol()) (src/ArmSlamCalib.cpp:490-493).
// Only the extrinsic K0 can be queried: that is the one marginal the reference computes and draws (:976-998).
class Marginals {
public:
    Marginals(const NonlinearFactorGraph& graph, const Values& solution) : impl_(graph, solution, ASC_BATCH_LM, default_params()) {}
    Matrix marginalCovariance(Key variable) {
        if (variable != impl_.extrinsicKey()) throw std::invalid_argument("Marginals: only the extrinsic " + DefaultKeyFormatter(impl_.extrinsicKey()) + " is supported");
        return impl_.extrinsicMarginal();
    }
private:
    static asc_params default_params() { asc_params a; asc_default_params(&a); return a; }
    facade_detail::BatchOptimizer impl_;
};

class DoglegOptimizer {
public:
    DoglegOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues, const DoglegParams& params = DoglegParams())
        : impl_(graph, initialValues, ASC_BATCH_DOGLEG, make(params)) {}
    const Values& optimize() { return impl_.optimize(); }
    double error() const { return impl_.error(); }
    int iterations() const { return impl_.iterations(); }
    const Values& values() const { return impl_.values(); }
private:
    static asc_params make(const DoglegParams& p) { asc_params a = facade_detail::base_params(p); a.delta_initial = p.deltaInitial; return a; }
    facade_detail::BatchOptimizer impl_;
};

}  // namespace gtsam
