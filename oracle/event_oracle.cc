// oracle/event_oracle.cc — CPU ORACLE (test infrastructure, never linked into the product).
//
// Restates reference src/Event/EventConversion.cc:
//   resolvePolarity :26-30, resolveMinMaxVals :32-39, roundFloatCoord :46-49, breakFloatCoords :51-57,
//   exp_XY2f :59-65, normalizeImage :67-72, ev2im :173-212, ev2im_gauss :215-269,
//   ev2mci_gg_f(Tcw,medDepth) :279-360, ev2mci_gg_f(params2D) :362-448,
// with Pinhole::project/unproject (src/CameraModels/Pinhole.cpp:30-62), MyCalibrator::isInImage
// (src/Utils/MyCalibrator.cpp:36-39) and Eigen::AngleAxisd(R) / AngleAxisd::toRotationMatrix (Eigen 3.3
// Geometry: matrix -> quaternion (Shepperd) -> angle/axis; Rodrigues) restated inline.
// Sums are accumulated sequentially in event order in float32, exactly like the reference; the running
// min/max over every intermediate sum (:257) is reproduced.  Pin for the degenerate DT==0 window (single
// timestamp): warp rate r := 0 (the reference computes 0*inf = NaN).
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

inline float resolvePolarity(bool withPol, bool evPol) { return (withPol && !evPol) ? -1.0f : 1.0f; }

inline void resolveMinMax(float v, float& mn, float& mx) {
    if (v > mx) mx = v;
    if (v < mn) mn = v;
}

inline bool inImage(float x, float y, int w, int h) { return (x >= 0 && x < float(w)) && (y >= 0 && y < float(h)); }

inline void breakFloatCoords(float X, float Y, int& x, int& y, float& xr, float& yr) {
    x = static_cast<int>(std::floor(X)); xr = X - float(x);
    y = static_cast<int>(std::floor(Y)); yr = Y - float(y);
}

inline float expXY2f(float x, float y, float sig2) {
    float dd = powf(x, 2) + powf(y, 2);
    dd /= 2.0f * sig2;
    return expf(-dd) / (2.0f * float(M_PI) * sig2);
}

struct Splat {
    float* img; int w, h; float sig2; int half; bool pol;
    float mn = 0.0f, mx = -1000000.0f;
    void add(float X, float Y, bool p) {
        int xi, yi; float xr, yr;
        breakFloatCoords(X, Y, xi, yi, xr, yr);
        for (int i = -half; i <= half; i++)
            for (int j = -half; j <= half; j++) {
                int xn = xi + i, yn = yi + j;
                float val = expXY2f(i - xr, j - yr, sig2);
                if (!inImage((float)xn, (float)yn, w, h)) continue;
                float nv = img[(size_t)yn * w + xn] + resolvePolarity(pol, p) * val;
                img[(size_t)yn * w + xn] = nv;
                resolveMinMax(nv, mn, mx);
            }
    }
};

// Eigen::AngleAxisd(const Matrix3d&)
void angleAxisFromR(const double R[3][3], double& angle, double axis[3]) {
    double q[4];  // x y z w
    double t = R[0][0] + R[1][1] + R[2][2];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[2][1] - R[1][2]) * t;
        q[1] = (R[0][2] - R[2][0]) * t;
        q[2] = (R[1][0] - R[0][1]) * t;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[k][j] - R[j][k]) * t;
        q[j] = (R[j][i] + R[i][j]) * t;
        q[k] = (R[k][i] + R[i][k]) * t;
    }
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (n != 0.0) {
        angle = 2.0 * std::atan2(n, std::fabs(q[3]));
        if (q[3] < 0) n = -n;
        axis[0] = q[0] / n; axis[1] = q[1] / n; axis[2] = q[2] / n;
    } else {
        angle = 0; axis[0] = 1; axis[1] = 0; axis[2] = 0;
    }
}

// Eigen::AngleAxisd::toRotationMatrix()
void rotFromAngleAxis(double angle, const double a[3], double R[3][3]) {
    double s = std::sin(angle), c = std::cos(angle);
    double sa[3] = {s * a[0], s * a[1], s * a[2]};
    double ca[3] = {(1 - c) * a[0], (1 - c) * a[1], (1 - c) * a[2]};
    double tmp;
    tmp = ca[0] * a[1]; R[0][1] = tmp - sa[2]; R[1][0] = tmp + sa[2];
    tmp = ca[0] * a[2]; R[0][2] = tmp + sa[1]; R[2][0] = tmp - sa[1];
    tmp = ca[1] * a[2]; R[1][2] = tmp - sa[0]; R[2][1] = tmp + sa[0];
    R[0][0] = ca[0] * a[0] + c; R[1][1] = ca[1] * a[1] + c; R[2][2] = ca[2] * a[2] + c;
}

// ---- camera models of the motion-compensation warp.  Pinhole: src/CameraModels/Pinhole.cpp:30-62.  KannalaBrandt8 (the camera of
// Examples/Event/EvMVSEC.yaml:50): src/CameraModels/KannalaBrandt8.cpp:86-103 (project, cv::Point3f), :111-129 (project, Eigen), :163-190
// (unproject: ten Newton steps on theta in float, precision = KB8_DEF_PRECISION = 1e-6, include/CameraModels/KannalaBrandt8.h:35).
// cam[0..3] = fx, fy, cx, cy; cam[4..7] = k1..k4 (KannalaBrandt8 only).
struct CamModel {
    int model; const float* p;
    void unproject(float x, float y, float& X, float& Y, float& Z) const {
        if (model == 0) { X = (x - p[2]) / p[0]; Y = (y - p[3]) / p[1]; Z = 1.f; return; }
        const float pwx = (x - p[2]) / p[0], pwy = (y - p[3]) / p[1];
        float scale = 1.f;
        float theta_d = sqrtf(pwx * pwx + pwy * pwy);
        theta_d = fminf(fmaxf(-3.1415926535897932384626433832795 / 2.f, theta_d), 3.1415926535897932384626433832795 / 2.f);
        if (theta_d > 1e-8) {
            float theta = theta_d;
            for (int j = 0; j < 10; j++) {
                float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
                float k0_theta2 = p[4] * theta2, k1_theta4 = p[5] * theta4;
                float k2_theta6 = p[6] * theta6, k3_theta8 = p[7] * theta8;
                float theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                                  (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
                theta = theta - theta_fix;
                if (fabsf(theta_fix) < 1e-6f) break;
            }
            scale = std::tan(theta) / theta_d;
        }
        X = pwx * scale; Y = pwy * scale; Z = 1.f;
    }
    // Eigen::Vector2d project(const Eigen::Vector3d&)
    void projectD(const double v[3], double& u, double& w) const {
        if (model == 0) { u = p[0] * v[0] / v[2] + p[2]; w = p[1] * v[1] / v[2] + p[3]; return; }
        const double x2_plus_y2 = v[0] * v[0] + v[1] * v[1];
        const double theta = atan2f(sqrtf(x2_plus_y2), v[2]);
        const double psi = atan2f(v[1], v[0]);
        const double theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2, theta7 = theta5 * theta2, theta9 = theta7 * theta2;
        const double r = theta + p[4] * theta3 + p[5] * theta5 + p[6] * theta7 + p[7] * theta9;
        u = p[0] * r * cos(psi) + p[2];
        w = p[1] * r * sin(psi) + p[3];
    }
    // cv::Point2f project(const cv::Point3f&)
    void projectF(float x, float y, float z, float& u, float& w) const {
        if (model == 0) { u = p[0] * x / z + p[2]; w = p[1] * y / z + p[3]; return; }
        const float x2_plus_y2 = x * x + y * y;
        const float theta = atan2f(sqrtf(x2_plus_y2), z);
        const float psi = atan2f(y, x);
        const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2, theta7 = theta5 * theta2, theta9 = theta7 * theta2;
        const float r = theta + p[4] * theta3 + p[5] * theta5 + p[6] * theta7 + p[7] * theta9;
        u = (float)(p[0] * r * cos(psi) + p[2]);
        w = (float)(p[1] * r * sin(psi) + p[3]);
    }
};

}  // namespace

extern "C" {

float orc_image_focus(const float* img, int w, int h, int patch, int what, int avg);

int orc_ev_accumulate(const orc_event* evs, int64_t n, int w, int h, float sigma, int mode, const float* Tcw16,
                      float depth, const float* K4, const float* se2, int se2_n, int pol, int normalize,
                      float* img, float* minmax2) {
    return orc_ev_accumulate_cam(evs, n, w, h, sigma, mode, Tcw16, depth, K4, 0, se2, se2_n, pol, normalize, img, minmax2);
}

int orc_ev_accumulate_cam(const orc_event* evs, int64_t n, int w, int h, float sigma, int mode, const float* Tcw16,
                          float depth, const float* cam8, int cam_model, const float* se2, int se2_n, int pol, int normalize,
                          float* img, float* minmax2) {
    const CamModel cam{cam_model, cam8};
    std::memset(img, 0, sizeof(float) * (size_t)w * h);
    float mn = 0.0f, mx = -1000000.0f;
    if (mode == 0) {
        for (int64_t i = 0; i < n; i++) {
            float ps = resolvePolarity(pol != 0, evs[i].p != 0);
            int px = static_cast<int>(roundf(evs[i].x)), py = static_cast<int>(roundf(evs[i].y));
            if (!inImage((float)px, (float)py, w, h)) continue;
            float nv = img[(size_t)py * w + px] + (ps * 0.001f);
            img[(size_t)py * w + px] = nv;
            resolveMinMax(nv, mn, mx);
        }
        if (minmax2) { minmax2[0] = mn; minmax2[1] = mx; }
        return (normalize && mx > mn) ? 1 : 0;   // 1: caller applies normalizeImage
    }
    Splat sp;
    sp.img = img; sp.w = w; sp.h = h; sp.sig2 = powf(sigma, 2);
    sp.half = static_cast<int>(std::ceil(sigma * 3.0)); sp.pol = pol != 0;
    if (mode == 1) {
        for (int64_t k = 0; k < n; k++) sp.add(evs[k].x, evs[k].y, evs[k].p != 0);
    } else if (mode == 2) {
        if (n <= 0) { if (minmax2) { minmax2[0] = mn; minmax2[1] = mx; } return 0; }
        double R[3][3], tt[3];
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R[r][c] = (double)Tcw16[r * 4 + c]; tt[r] = (double)Tcw16[r * 4 + 3]; }
        double ang, ax[3];
        angleAxisFromR(R, ang, ax);
        double t1 = evs[n - 1].ts, DT = t1 - evs[0].ts, invDT = 1.0 / DT;
        for (int64_t k = 0; k < n; k++) {
            double rate = DT > 0 ? (t1 - evs[k].ts) * invDT : 0.0;
            float X, Y, Z;
            cam.unproject(evs[k].x, evs[k].y, X, Y, Z);
            double P[3] = {(double)X, (double)Y, (double)Z};
            double nR[3][3];
            rotFromAngleAxis(ang * rate, ax, nR);
            double np[3];
            for (int r = 0; r < 3; r++) {
                double md[3] = {(double)depth * nR[r][0], (double)depth * nR[r][1], (double)depth * nR[r][2]};
                np[r] = (md[0] * P[0] + md[1] * P[1] + md[2] * P[2]) + tt[r] * rate;
            }
            double u, v;
            cam.projectD(np, u, v);
            sp.add((float)u, (float)v, evs[k].p != 0);
        }
    } else if (mode == 3) {
        if (n <= 0) { if (minmax2) { minmax2[0] = mn; minmax2[1] = mx; } return 0; }
        double t1 = evs[n - 1].ts;
        float DT = static_cast<float>(t1 - evs[0].ts);
        float invDT = 1.f / DT;
        float omega0 = se2[0] * invDT, vx0 = se2[1] * invDT, vy0 = se2[2] * invDT;
        float sc = 1.f;
        if (se2_n > 3) sc = se2[3];
        float scDiff = 1.f - sc;
        for (int64_t k = 0; k < n; k++) {
            float tk = static_cast<float>(t1 - evs[k].ts);
            float X, Y, Z;
            cam.unproject(evs[k].x, evs[k].y, X, Y, Z);
            float th = tk * omega0;
            float cs = scDiff * (1 - tk * invDT) + sc;
            float xp = cs * (X * cosf(th) - Y * sinf(th)) + vx0 * tk;
            float yp = cs * (X * sinf(th) + Y * cosf(th)) + vy0 * tk;
            float u, v;
            cam.projectF(xp, yp, Z, u, v);
            sp.add(u, v, evs[k].p != 0);
        }
    } else {
        return -3;
    }
    if (minmax2) { minmax2[0] = sp.mn; minmax2[1] = sp.mx; }
    return normalize ? 1 : 0;
}

void orc_normalize_convert_u8(const float* img, int n, float maxVal, float minVal, uint8_t* out) {
    float alpha = 255.f / (maxVal - minVal);
    float beta = -minVal * alpha;
    for (int i = 0; i < n; i++) {
        float v = img[i] * alpha + beta;
        int r = (int)lrintf(v);
        out[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

void orc_normalize_minmax_u8(const float* img, int n, uint8_t* out) {
    if (n <= 0) return;
    double smin = img[0], smax = img[0];
    for (int i = 1; i < n; i++) { if (img[i] < smin) smin = img[i]; if (img[i] > smax) smax = img[i]; }
    double scale = 255.0 * (smax - smin > 2.220446049250313e-16 ? 1. / (smax - smin) : 0);
    double shift = 0.0 - smin * scale;
    float a = (float)scale, b = (float)shift;
    for (int i = 0; i < n; i++) {
        float v = img[i] * a + b;
        int r = (int)lrintf(v);
        out[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

// ---- Jacobian of the contrast objective w.r.t. the window's SE3 motion: SURVEY.md §8f rank 2 (second half)
// EvImConverter::ev2mci_gg_f_jac, src/Event/EventConversion.cc:533-662 (called once per optimiser iteration by
// EvMciEdge::linearizeOplus, src/Utils/MyOptimTypes.cpp:16).  Rt12 = rotation (row-major 3x3) then translation of the
// vertex estimate, in double as the reference reads them from g2o.  Seven float images (I and dI/d[wx wy wz vx vy vz])
// are splatted, jac[k] = -2 * imageMean(I .* I_k, global): cv::mean (global) or the mean of 30x30-cell means.
void orc_ev_mci_jac(const orc_event* evs, int64_t n, int w, int h, float sigma, const double* Rt12, float medDepth, const float* K4, int pol,
                    int global, double* jac6) {
    for (int k = 0; k < 6; k++) jac6[k] = 0.0;
    if (n <= 0) return;
    const float sig2 = powf(sigma, 2), invSig2 = 1.f / sig2;
    const int half = static_cast<int>(ceil(sigma * 3.0));
    const size_t npx = (size_t)w * h;
    std::vector<float> im(7 * npx, 0.f);
    double R[3][3], ang, ax[3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r][c] = Rt12[3 * r + c];
    const double t[3] = {Rt12[9], Rt12[10], Rt12[11]};
    angleAxisFromR(R, ang, ax);
    const double t1 = evs[n - 1].ts, DT = t1 - evs[0].ts, invDT = 1.0 / DT;
    const float fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    for (int64_t e = 0; e < n; e++) {
        const float ex = evs[e].x, ey = evs[e].y;
        const double rate = (t1 - evs[e].ts) * invDT;
        const float ux = (ex - cx) / fx, uy = (ey - cy) / fy;              // Pinhole::unproject (float)
        const double X3[3] = {(double)medDepth * (double)ux, (double)medDepth * (double)uy, (double)medDepth * 1.0};
        double Rk[3][3];
        rotFromAngleAxis(ang * rate, ax, Rk);
        double P[3];
        for (int r = 0; r < 3; r++) P[r] = (Rk[r][0] * X3[0] + Rk[r][1] * X3[1] + Rk[r][2] * X3[2]) + t[r] * rate;
        const double X = P[0], Y = P[1], Z = P[2];
        double S[3][6] = {{0, Z, -Y, 1, 0, 0}, {-Z, 0, X, 0, 1, 0}, {Y, -X, 0, 0, 0, 1}};
        for (int r = 0; r < 3; r++) for (int c = 0; c < 6; c++) S[r][c] *= rate;
        const double J00 = (double)fx / Z, J02 = -(double)fx * X / (Z * Z), J11 = (double)fy / Z, J12 = -(double)fy * Y / (Z * Z);
        double JP[2][6];
        for (int c = 0; c < 6; c++) {
            JP[0][c] = -(J00 * S[0][c] + 0.0 * S[1][c] + J02 * S[2][c]);
            JP[1][c] = -(0.0 * S[0][c] + J11 * S[1][c] + J12 * S[2][c]);
        }
        const double u = (double)fx * X / Z + (double)cx, v = (double)fy * Y / Z + (double)cy;   // Pinhole::project
        int xi, yi; float xr, yr;
        breakFloatCoords((float)u, (float)v, xi, yi, xr, yr);
        const float ps = resolvePolarity(pol != 0, evs[e].p != 0);
        for (int i = -half; i <= half; i++)
            for (int j = -half; j <= half; j++) {
                const int xn = xi + i, yn = yi + j;
                if (!inImage((float)xn, (float)yn, w, h)) continue;
                const float val = expXY2f(i - xr, j - yr, sig2);
                const float gx = invSig2 * (i - xr) * val, gy = invSig2 * (j - yr) * val;
                const size_t o = (size_t)yn * w + xn;
                im[o] = im[o] + ps * val;
                for (int k = 0; k < 6; k++) {
                    const double JI = (double)gx * JP[0][k] + (double)gy * JP[1][k];
                    im[(size_t)(k + 1) * npx + o] = (float)((double)im[(size_t)(k + 1) * npx + o] + (double)ps * JI);
                }
            }
    }
    std::vector<float> prod(npx);
    for (int k = 0; k < 6; k++) {
        for (size_t o = 0; o < npx; o++) prod[o] = im[o] * im[(size_t)(k + 1) * npx + o];
        float m;
        if (global) { double s = 0; for (size_t o = 0; o < npx; o++) s += prod[o]; m = (float)(s / (double)npx); }
        else m = orc_image_focus(prod.data(), w, h, 30, 2, 1);
        jac6[k] = (double)(-m * 2);
    }
}

// ---- contrast metric of a (motion-compensated) event frame: SURVEY.md §8f rank 2
// EvImConverter::measureImageFocusLocal / measureImageFocusGlobal / imageMeanLocal, src/Event/EventConversion.cc:79-162,
// DEF_PATCH_SIZE_STD = 30 (include/Event/EventConversion.h:28).  cv::meanStdDev on CV_32F accumulates sum and
// sum of squares in double; mean = s/N, std = sqrt(max(sq/N - mean^2, 0)).  The per-patch values are cast to float
// and accumulated in float in patch order (:98-101), the median variant sorts them (:108-109).
// what: 0 = local std-dev (measureImageFocusLocal), 1 = global std-dev (measureImageFocusGlobal), 2 = local mean (imageMeanLocal)
// avg:  1 = average of the patch values, 0 = median (vLocalStd[cnt/2] after sorting)
static void patch_mean_std(const float* img, int w, int r0, int r1, int c0, int c1, double& mean, double& sd) {
    double s = 0, sq = 0;
    for (int y = r0; y < r1; y++)
        for (int x = c0; x < c1; x++) { const double v = img[(size_t)y * w + x]; s += v; sq += v * v; }
    const double n = (double)(r1 - r0) * (c1 - c0);
    mean = s / n;
    const double var = sq / n - mean * mean;
    sd = sqrt(var > 0 ? var : 0);
}

float orc_image_focus(const float* img, int w, int h, int patch, int what, int avg) {
    if (what == 1) { double m, sd; patch_mean_std(img, w, 0, h, 0, w, m, sd); return (float)sd; }
    float acc = 0.f;
    int cnt = 0;
    float vals[4096];
    for (int i = 0; i < h; i += patch)
        for (int j = 0; j < w; j += patch) {
            double m, sd;
            patch_mean_std(img, w, i, i + patch < h ? i + patch : h, j, j + patch < w ? j + patch : w, m, sd);
            const float v = (float)(what == 2 ? m : sd);
            acc += v;
            if (cnt < 4096) vals[cnt] = v;
            cnt++;
        }
    if (avg) return acc / (float)cnt;
    const int n = cnt < 4096 ? cnt : 4096;
    for (int a = 1; a < n; a++) { const float v = vals[a]; int b = a - 1; while (b >= 0 && vals[b] > v) { vals[b + 1] = vals[b]; b--; } vals[b + 1] = v; }
    return vals[cnt / 2];
}

}  // extern "C"
