// Builds the simulated calibration graph the way ArmSlamCalib does (AddFactor / AddValue funnels, src/ArmSlamCalib.cpp
// :299-327, :654-706, :1625-1651, :1916-1941) on top of the GTSAM-shaped facade, then runs the OptimizeStep body
// (:438-467) unchanged for the chosen optimization_mode.  Prints "RESULT <iterations> <error> <x y z of K0>".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>

#include "gtsam_facade/gtsam_facade.h"

using namespace gtsam;

static Key ExtrinsicSymbol() { return Symbol('K', 0); }          // ArmSlamCalib.h:284
static Key ConfigSymbol(size_t i) { return Symbol('q', i); }     // :285
static Key LandmarkSymbol(size_t i) { return Symbol('l', i); }   // :286

enum OptimizationMode { BatchLM, BatchDogleg, ISAM };            // ArmSlamCalib.h:59-64

int main(int argc, char** argv) {
    const std::string mode_name = argc > 1 ? argv[1] : "BatchLM";
    const int T = argc > 2 ? atoi(argv[2]) : 40;
    OptimizationMode optimizationMode = BatchDogleg;              // Params() default, ArmSlamCalib.h:97
    if (mode_name == "BatchLM") optimizationMode = BatchLM;       // string parse, ArmSlamCalib.cpp:415-422
    else if (mode_name == "BatchDogLeg" || mode_name == "BatchDogleg") optimizationMode = BatchDogleg;

    // ---- simulated problem from the host generator (stands in for CreateSimulation + SimulationStep)
    double t_p2j[6 * 12], axis[6 * 3], t_c2j[6 * 12];
    asc_chain_mico(t_p2j, axis, t_c2j);
    asc_sim_config cfg;
    asc_sim_default_config(&cfg);
    cfg.trajectory_size = T;
    cfg.seed = 3;
    asc_srand(3);
    asc_sim_problem* sim = nullptr;
    if (asc_sim_create(&cfg, 6, t_p2j, axis, t_c2j, 640, 640, 320, 240, &sim) != ASC_OK) { fprintf(stderr, "sim failed\n"); return 2; }
    int32_t P = 0, L = 0, Ls = 0; int64_t M = 0;
    asc_sim_sizes(sim, &P, &L, &M, &Ls);
    std::vector<int32_t> pose_idx(M), lm_idx(M), lm_id(L);
    std::vector<double> uv(2 * M), enc(6 * P), qt(6 * P), lt(3 * L);
    std::vector<int64_t> order(M);
    asc_sim_export(sim, pose_idx.data(), lm_idx.data(), uv.data(), enc.data(), qt.data(), lt.data(), lm_id.data(), order.data());
    asc_sim_destroy(sim);

    asc_facade::SerialChainPtr arm(new asc_facade::SerialChain());
    arm->t_p2j.assign(t_p2j, t_p2j + 72); arm->axis.assign(axis, axis + 18); arm->t_c2j.assign(t_c2j, t_c2j + 72);

    // ---- the graph, built like the reference builds it
    std::shared_ptr<NonlinearFactorGraph> graph = std::make_shared<NonlinearFactorGraph>();
    Values currentEstimate;
    std::shared_ptr<Cal3_S2> calib = std::make_shared<Cal3_S2>(640, 640, 0.0, 320, 240);                        // :670
    noiseModel::mEstimator::Cauchy::shared_ptr cauchyEstimator = std::make_shared<noiseModel::mEstimator::Cauchy>(3.0);   // :197-199
    SharedNoiseModel encoderNoise = noiseModel::Diagonal::Sigmas(Vector::Constant(6, 0.1));                     // :672-677
    SharedNoiseModel driftNoise = noiseModel::Diagonal::Sigmas(Vector::Constant(6, 0.05));                       // :678, drift_noise_level
    const bool addDriftNoise = argc > 3 && std::string(argv[3]) == "drift";                                     // use_drift_noise, :373
    SharedNoiseModel calibrationPrior = noiseModel::Diagonal::Sigmas(Vector::Constant(6, 0.1));                 // :684
    SharedNoiseModel measurementNoise = noiseModel::Robust::Create(cauchyEstimator, noiseModel::Diagonal::Sigmas(Vector::Constant(2, 1.0)));   // :704
    SharedNoiseModel landmarkPrior = noiseModel::Robust::Create(cauchyEstimator, noiseModel::Diagonal::Sigmas(Vector::Constant(3, 50.0)));     // :705
    graph->add(std::make_shared<PriorFactor<Pose3> >(ExtrinsicSymbol(), Pose3::identity(), calibrationPrior));   // :685
    currentEstimate.insert(ExtrinsicSymbol(), Pose3::identity());                                                // :690
    for (int i = 0; i < P; ++i) {                                                                                // SimulationStep :307-327
        Vector e(6);
        for (int k = 0; k < 6; ++k) e(k) = enc[6 * i + k];
        currentEstimate.insert(ConfigSymbol(i), RobotConfig(e, arm));
        graph->add(std::make_shared<EncoderFactor>(ConfigSymbol(i), e, encoderNoise, false, 0.05));
        if (addDriftNoise && i > 0) {                                                                            // :317-320 -> AddDriftFactor :1653-1668
            Vector e0(6);
            for (int k = 0; k < 6; ++k) e0(k) = enc[6 * (i - 1) + k];
            graph->add(std::make_shared<DriftFactor>(ConfigSymbol(i - 1), ConfigSymbol(i), e0, e, driftNoise));
        }
    }
    for (int j = 0; j < L; ++j) {                                                                                // :1869-1870
        const Point3 pos(lt[3 * j], lt[3 * j + 1], lt[3 * j + 2]);
        currentEstimate.insert(LandmarkSymbol(lm_id[j]), pos);
        graph->add(std::make_shared<PriorFactor<Point3> >(LandmarkSymbol(lm_id[j]), pos, landmarkPrior));
    }
    for (int64_t m = 0; m < M; ++m)                                                                              // AddObservationFactor :1625-1646
        graph->add(std::make_shared<RobotProjectionFactor<Cal3_S2> >(Point2(uv[2 * m], uv[2 * m + 1]), measurementNoise, ConfigSymbol(pose_idx[m]),
                                                                    LandmarkSymbol(lm_id[lm_idx[m]]), ExtrinsicSymbol(), arm, nullptr, nullptr, calib,
                                                                    (size_t)pose_idx[m], (size_t)lm_id[lm_idx[m]], false, true));

    // ---- ArmSlamCalib::OptimizeStep, batch branches (src/ArmSlamCalib.cpp:438-467), source unchanged
    double total_error = 0;
    int iterations = 0;
    try {
        switch (optimizationMode) {
            case BatchDogleg: {
                DoglegParams doglegParams;
                DoglegOptimizer optimizer(*graph, currentEstimate, doglegParams);
                try {
                    currentEstimate = optimizer.optimize();
                    total_error = optimizer.error();
                    iterations = optimizer.iterations();
                } catch (gtsam::IndeterminantLinearSystemException& e) {
                    const gtsam::KeyFormatter keyFormatter = gtsam::DefaultKeyFormatter;
                    fprintf(stderr, "Indeterminate: %s\n", keyFormatter(e.nearbyVariable()).c_str());
                }
                break;
            }
            case BatchLM: {
                LevenbergMarquardtParams params;
                params.linearSolverType = LevenbergMarquardtParams::SEQUENTIAL_CHOLESKY;
                LevenbergMarquardtOptimizer optimizer(*graph, currentEstimate, params);
                currentEstimate = optimizer.optimize();
                total_error = optimizer.error();
                iterations = optimizer.iterations();
                break;
            }
            case ISAM: break;
        }
    } catch (std::exception& e) {
        fprintf(stderr, "EXCEPTION %s\n", e.what());
        return 3;
    }
    const Pose3& X = currentEstimate.at<Pose3>(ExtrinsicSymbol());
    // copy landmark estimates back like :522-529
    size_t found = 0;
    for (int j = 0; j < L; ++j)
        if (currentEstimate.find(LandmarkSymbol(lm_id[j])) != currentEstimate.end()) { (void)currentEstimate.at<gtsam::Point3>(LandmarkSymbol(lm_id[j])); ++found; }
    printf("RESULT %d %.12g %.12g %.12g %.12g %zu\n", iterations, total_error, X.translation().x(), X.translation().y(), X.translation().z(), found);
    if (argc > 3 && std::string(argv[3]) == "dot") {
        // the rest of the GTSAM surface ArmSlamCalib.cpp touches: saveGraph (:1910-1913), clone (:2209), Rot3::toQuaternion (:2224), Point3 ops (:1377)
        std::ostringstream os;
        graph->saveGraph(os, currentEstimate);
        NonlinearFactorGraph copy = graph->clone();
        const Quaternion qx = X.rotation().toQuaternion();
        const double d = (currentEstimate.at<Point3>(LandmarkSymbol(lm_id[0])) - Point3(lt[0], lt[1], lt[2])).norm();
        printf("DOT %zu %zu %.6f %.6f\n", os.str().size(), copy.size(), qx.w() * qx.w() + qx.x() * qx.x() + qx.y() * qx.y() + qx.z() * qx.z(), d);
    }
    if (argc > 3 && std::string(argv[3]) == "marginal") {
        // batch counterpart of  extrinsicMarginals = isam2.marginalCovariance(ExtrinsicSymbol());  (src/ArmSlamCalib.cpp:490-493)
        try {
            Marginals marginals(*graph, currentEstimate);
            Matrix extrinsicMarginals = marginals.marginalCovariance(ExtrinsicSymbol());
            if (extrinsicMarginals.rows() > 0)   // the guard of DrawState, :976
                printf("MARGINAL %.12g %.12g %.12g %.12g %.12g %.12g\n", extrinsicMarginals(0, 0), extrinsicMarginals(1, 1), extrinsicMarginals(2, 2),
                       extrinsicMarginals(3, 3), extrinsicMarginals(4, 4), extrinsicMarginals(5, 5));
        } catch (std::exception& e) {
            fprintf(stderr, "EXCEPTION %s\n", e.what());
            return 3;
        }
    }
    return 0;
}
