// ps2 — the YAML-configured Problem Set 2 executable on the B200 library.
//
// Same surface as the reference's binary (ProblemSets/ps2_cpp/src/main.cpp:329-361): reads the keys of
// config/ps2.yaml (images, output_dir, use_gpu_disparity, problem_N_{ssd,ncorr}.{window_radius,
// disparity_range}), runs the five problems, writes ps2-<problem>-<part>-<n>[-inverted].png into
// output_dir and the same log lines to stdout and ps2.log.  The disparity maps come from
// cuda::disparitySSD / cuda::disparityNCorr of include/stereo_b200.hpp, called exactly like
// disparitySSDPair / disparityNCorrPair do (main.cpp:21-78).
//
//   ps2 [config.yaml]        default ../config/ps2.yaml, the reference's fixed path (main.cpp:16)
//   ps2 --selftest <dir>     host-side checks (YAML, PNG, preprocessing) without a GPU; writes files
//                            that tests/test_ps2_app.py compares with executed OpenCV
//
// Differences, all deliberate: `use_gpu_disparity: false` is an error (this build has no CPU path — the
// reference's serial:: functions are the *oracle* here, not product code); optional keys `device`
// (GPU ordinal, default 0), `noise_seed` (cv::RNG state, default OpenCV's 0xffffffff),
// `opencv_gray_shift` (14 = the RGB2GRAY coefficients of OpenCV 3.4.1, the default; 15 = OpenCV >= 3.4.2) and
// `gpus` (default 1 = the reference's behaviour; N > 1: every pair's output rows are sharded over the first N B200s in
// row bands with halo, stereo_mgpu_*, same maps).
// Both maps of a pair come from ONE library call (one upload of the pair; SSD pairs from one cost volume), so each of the
// two "... execution took" lines the reference prints per pair reports half of that call's device time.
#include "imgproc.hpp"
#include "yaml_subset.hpp"

#include <sys/stat.h>

#include <chrono>
#include <cstdarg>
#include <ctime>
#include <functional>
#include <memory>

namespace {

// ---- logging: spdlog's default pattern, two loggers sharing ps2.log (main.cpp:331-340) -------------------
FILE* g_logfile = nullptr;
void log_line(bool to_stdout, const char* logger, const char* level, const char* fmt, va_list ap) {
    char msg[1024];
    std::vsnprintf(msg, sizeof(msg), fmt, ap);
    const auto now = std::chrono::system_clock::now();
    const std::time_t t = std::chrono::system_clock::to_time_t(now);
    const int ms = int(std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
    std::tm tmv; localtime_r(&t, &tmv);
    char stamp[64];
    std::strftime(stamp, sizeof(stamp), "%Y-%m-%d %H:%M:%S", &tmv);
    if (to_stdout) { std::printf("[%s.%03d] [%s] [%s] %s\n", stamp, ms, logger, level, msg); std::fflush(stdout); }
    if (g_logfile) { std::fprintf(g_logfile, "[%s.%03d] [%s] [%s] %s\n", stamp, ms, logger, level, msg); std::fflush(g_logfile); }
}
void info(const char* fmt, ...) { va_list ap; va_start(ap, fmt); log_line(true, "logger", "info", fmt, ap); va_end(ap); }
void warn(const char* fmt, ...) { va_list ap; va_start(ap, fmt); log_line(true, "logger", "warning", fmt, ap); va_end(ap); }
void error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); log_line(true, "logger", "error", fmt, ap); va_end(ap); }
void file_info(const char* fmt, ...) { va_list ap; va_start(ap, fmt); log_line(false, "file_logger", "info", fmt, ap); va_end(ap); }

struct DisparityParams { size_t window_radius = 0; int disparity_range = 0; };   // Config::DisparitySSD (include/Config.h:40-46)

struct Config {
    std::map<std::string, ps2::Image8> images;
    std::string output_dir = "./";
    bool use_gpu = false;                                    // Config.h:49: defaults to false
    DisparityParams p[6];
    uint64_t noise_seed = 0xffffffffu;
    int device = 0;
    int gpus = 1;                                            // > 1: row-band sharding over the first `gpus` devices
    int gray_shift = 14;                                     // OpenCV 3.4.1's RGB2GRAY coefficients (15: OpenCV >= 3.4.2)
};

bool make_dir(const std::string& path) {                      // common::makeDir (common/src/Utils.cpp:17-24)
    if (mkdir(path.c_str(), 0775) == 0) return true;
    struct stat st;
    return errno == EEXIST && stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

Config load_config(const std::string& path) {
    ps2::YamlDoc y = ps2::YamlDoc::load(path);
    Config cfg;
    static const char* image_keys[] = {"pair0-L", "pair0-R", "pair1-L", "pair1-R", "pair2-L", "pair2-R"};
    for (const char* k : image_keys) {                         // Config::Images (lib/Config.cpp:8-17); ground-truth maps are loaded
        const std::string key = std::string("images.") + k;   // by the reference but never used, so they are optional here
        if (!y.has(key)) throw std::runtime_error("Loading images failed! (missing " + key + ")");
        cfg.images[k] = ps2::imread(y.str(key));
        info("Loaded %s from %s (%d x %d, %d ch)", k, y.str(key).c_str(), cfg.images[k].cols, cfg.images[k].rows, cfg.images[k].channels);
    }
    bool made = false;
    if (y.has("output_dir")) {
        cfg.output_dir = y.str("output_dir");
        if (make_dir(cfg.output_dir)) { info("Created output directory at \"%s\"", cfg.output_dir.c_str()); made = true; }
    }
    if (!made) { cfg.output_dir = "./"; warn("No output path specified or could not make new directory; using current directory"); }
    if (y.has("use_gpu_disparity")) {
        cfg.use_gpu = y.boolean("use_gpu_disparity");
        info("Using %s for disparity computation", cfg.use_gpu ? "GPU" : "CPU");
    }
    static const char* sections[6] = {nullptr, "problem_1_ssd", "problem_2_ssd", "problem_3_ssd", "problem_4_ncorr", "problem_5_ncorr"};
    for (int i = 1; i <= 5; ++i) {
        const std::string s = sections[i];
        if (!y.has(s + ".window_radius") || !y.has(s + ".disparity_range"))
            throw std::runtime_error("Loading Problem " + std::to_string(i) + " parameters failed!");
        cfg.p[i].window_radius = size_t(y.integer(s + ".window_radius"));
        cfg.p[i].disparity_range = int(y.integer(s + ".disparity_range"));
    }
    if (y.has("noise_seed")) cfg.noise_seed = uint64_t(y.integer("noise_seed"));
    if (y.has("device")) cfg.device = int(y.integer("device"));
    if (y.has("gpus")) cfg.gpus = int(y.integer("gpus"));
    if (y.has("opencv_gray_shift")) cfg.gray_shift = int(y.integer("opencv_gray_shift")) == 15 ? 15 : 14;
    return cfg;
}

// disparitySSDPair / disparityNCorrPair (main.cpp:21-78): left-referenced over [-range, 0], then
// right-referenced (images swapped) over [0, +range].
sb::MultiGpu* g_mgpu = nullptr;
void disparity_pair(int cost, const char* kernel_name, const sb::Mat& left, const sb::Mat& right, const DisparityParams& p,
                    sb::Mat& left_disp, sb::Mat& right_disp) {
    file_info("Setting up CUDA kernel execution...");                                  // DisparitySSD.cu:157-158,191,203
    file_info("Original image: rows=%d cols=%d", left.rows, left.cols);
    file_info("Launching %s", kernel_name);
    double ms;
    if (g_mgpu) {
        g_mgpu->disparityPairBands(cost, left, right, p.window_radius, p.disparity_range, left_disp, right_disp);
        ms = 0;
        for (int k = 0; k < g_mgpu->deviceCount(); ++k) { const double t = stereo_ctx_last_kernel_ms(stereo_mgpu_ctx(g_mgpu->handle(), k)); if (t > ms) ms = t; }
    } else {
        sb::run_pair(cost, left, right, p.window_radius, p.disparity_range, left_disp, right_disp, sb::S8C1, nullptr);
        ms = double(sb::lastKernelMs());
    }
    file_info("%s execution took %g ms", kernel_name, ms / 2);
    file_info("Setting up CUDA kernel execution...");
    file_info("Original image: rows=%d cols=%d", left.rows, left.cols);
    file_info("Launching %s", kernel_name);
    file_info("%s execution took %g ms", kernel_name, ms / 2);
}

void write_maps(const Config& cfg, const std::string& stem, sb::Mat& left_disp, sb::Mat& right_disp, bool with_inverted) {
    const sb::Mat l8 = ps2::normalize_minmax_u8(left_disp);                                // main.cpp:94-99, 123-132
    ps2::imwrite_gray(cfg.output_dir + "/" + stem + "-1.png", l8.data, l8.rows, l8.cols, l8.step);
    if (with_inverted) { const sb::Mat inv = ps2::inverted(l8); ps2::imwrite_gray(cfg.output_dir + "/" + stem + "-1-inverted.png", inv.data, inv.rows, inv.cols, inv.step); }
    const sb::Mat r8 = ps2::normalize_minmax_u8(right_disp);
    ps2::imwrite_gray(cfg.output_dir + "/" + stem + "-2.png", r8.data, r8.rows, r8.cols, r8.step);
}

struct Variant { char part; enum Kind { Clean, Noisy, Contrast } kind; };

// ---- device images (stereo_dev_* / stereo_image_*): a pair is uploaded once, as 8-bit pixels, and its grey float images
//      stay on the device for both directions and for every problem that uses the pair ------------------------------------
struct DevImage {
    void* p = nullptr; size_t step = 0; int rows = 0, cols = 0;
    DevImage() = default;
    DevImage(int r, int c, size_t elem) : rows(r), cols(c) {
        step = (size_t(c) * elem + 255) & ~size_t(255);
        sb::check(stereo_dev_alloc(sb::default_ctx(), step * r, &p), "stereo_dev_alloc");
    }
    DevImage(DevImage&& o) noexcept { *this = std::move(o); }
    DevImage& operator=(DevImage&& o) noexcept { std::swap(p, o.p); step = o.step; rows = o.rows; cols = o.cols; return *this; }
    DevImage(const DevImage&) = delete;
    DevImage& operator=(const DevImage&) = delete;
    ~DevImage() { if (p) stereo_dev_free(sb::default_ctx(), p); }
    float* f32() const { return static_cast<float*>(p); }
};
std::map<std::string, DevImage> g_gray;             // image key -> CV_32FC1 grey image on the device

const DevImage& device_gray(const Config& cfg, const std::string& key, bool gray_from_colour) {
    auto it = g_gray.find(key);
    if (it != g_gray.end()) return it->second;
    const ps2::Image8& img = cfg.images.at(key);
    if (!gray_from_colour && img.channels != 1) throw std::runtime_error("expected a single-channel image (the reference asserts CV_32FC1, main.cpp:27)");
    stereo_ctx* ctx = sb::default_ctx();
    DevImage raw(img.rows, img.cols * img.channels, 1), gray(img.rows, img.cols, 4);
    sb::check(stereo_dev_upload(ctx, raw.p, raw.step, img.data.data(), size_t(img.cols) * img.channels, size_t(img.cols) * img.channels, img.rows), "stereo_dev_upload");
    sb::check(stereo_image_gray_f32_device(ctx, static_cast<const uint8_t*>(raw.p), raw.step, img.rows, img.cols, img.channels, cfg.gray_shift,
                                           gray.f32(), gray.step, nullptr), "stereo_image_gray_f32_device");
    return g_gray.emplace(key, std::move(gray)).first->second;
}

// One problem = one image pair, one cost, and a list of input variants (main.cpp:80-327).
void run_problem(const Config& cfg, ps2::CvRng& rng, int number, const char* pair, bool gray_from_colour, int fn, const char* kernel_name,
                 std::initializer_list<Variant> variants, bool with_inverted) {
    info("Problem %d begins", number);
    const auto start = std::chrono::high_resolution_clock::now();
    const ps2::Image8& li = cfg.images.at(std::string(pair) + "-L");
    const ps2::Image8& ri = cfg.images.at(std::string(pair) + "-R");
    if (li.rows != ri.rows || li.cols != ri.cols) throw std::runtime_error("left/right image sizes differ (main.cpp:28)");
    sb::Mat ld, rd;
    if (g_mgpu) {         // several GPUs: host images, output rows sharded (stereo_mgpu_*)
        const sb::Mat left = gray_from_colour ? ps2::rgb2gray_on_bgr_as_float(li, cfg.gray_shift) : ps2::to_float(li);
        const sb::Mat right = gray_from_colour ? ps2::rgb2gray_on_bgr_as_float(ri, cfg.gray_shift) : ps2::to_float(ri);
        for (const Variant& v : variants) {
            sb::Mat l = left, r = right;
            if (v.kind == Variant::Noisy) ps2::add_noise(rng, left, right, 0.f, 10.f, l, r);   // main.cpp:169
            if (v.kind == Variant::Contrast) { l = ps2::scaled(left, 1.1f); r = ps2::scaled(right, 1.1f); }   // main.cpp:191-193
            disparity_pair(fn, kernel_name, l, r, cfg.p[number], ld, rd);
            write_maps(cfg, "ps2-" + std::to_string(number) + "-" + v.part, ld, rd, with_inverted);
        }
    } else {
        stereo_ctx* ctx = sb::default_ctx();
        const DevImage& gl = device_gray(cfg, std::string(pair) + "-L", gray_from_colour);
        const DevImage& gr = device_gray(cfg, std::string(pair) + "-R", gray_from_colour);
        const int rows = gl.rows, cols = gl.cols;
        DevImage vl(rows, cols, 4), vr(rows, cols, 4), nz(rows, cols, 4), dl(rows, cols, 1), dr(rows, cols, 1);
        for (const Variant& v : variants) {
            const float* l = gl.f32(); const float* r = gr.f32();
            if (v.kind == Variant::Noisy) {               // addNoise (main.cpp:140-153): cv::randn's stream on the host, the add on the device
                sb::Mat noise(rows, cols, sb::F32C1);
                const DevImage* src[2] = {&gl, &gr}; DevImage* dst[2] = {&vl, &vr};
                for (int k = 0; k < 2; ++k) {
                    rng.fill_normal(noise, 0.f, 10.f);
                    sb::check(stereo_dev_upload(ctx, nz.p, nz.step, noise.data, noise.step, size_t(cols) * 4, rows), "stereo_dev_upload");
                    sb::check(stereo_image_scale_add_f32_device(ctx, src[k]->f32(), src[k]->step, nz.f32(), nz.step, 1.f, rows, cols, dst[k]->f32(), dst[k]->step, nullptr),
                              "stereo_image_scale_add_f32_device");
                    sb::check(stereo_ctx_synchronize(ctx, nullptr), "stereo_ctx_synchronize");     // nz is reused for the second image
                }
                l = vl.f32(); r = vr.f32();
            } else if (v.kind == Variant::Contrast) {     // main.cpp:191-193
                sb::check(stereo_image_scale_add_f32_device(ctx, gl.f32(), gl.step, nullptr, 0, 1.1f, rows, cols, vl.f32(), vl.step, nullptr), "scale");
                sb::check(stereo_image_scale_add_f32_device(ctx, gr.f32(), gr.step, nullptr, 0, 1.1f, rows, cols, vr.f32(), vr.step, nullptr), "scale");
                l = vl.f32(); r = vr.f32();
            }
            const size_t step = v.kind == Variant::Clean ? gl.step : vl.step;
            file_info("Setting up CUDA kernel execution...");                                  // DisparitySSD.cu:157-158,191,203
            file_info("Original image: rows=%d cols=%d", rows, cols);
            file_info("Launching %s", kernel_name);
            sb::check(stereo_disparity_pair_f32_device(ctx, fn, l, step, r, step, rows, cols, int(cfg.p[number].window_radius), cfg.p[number].disparity_range,
                                                       dl.p, dr.p, dl.step, 1, nullptr), kernel_name);
            ld.create(rows, cols, sb::S8C1); rd.create(rows, cols, sb::S8C1);
            sb::check(stereo_dev_download(ctx, ld.data, ld.step, dl.p, dl.step, size_t(cols), rows), "stereo_dev_download");
            sb::check(stereo_dev_download(ctx, rd.data, rd.step, dr.p, dr.step, size_t(cols), rows), "stereo_dev_download");
            const double ms = double(sb::lastKernelMs());
            file_info("%s execution took %g ms", kernel_name, ms / 2);
            file_info("Setting up CUDA kernel execution...");
            file_info("Original image: rows=%d cols=%d", rows, cols);
            file_info("Launching %s", kernel_name);
            file_info("%s execution took %g ms", kernel_name, ms / 2);
            write_maps(cfg, "ps2-" + std::to_string(number) + "-" + v.part, ld, rd, with_inverted);
        }
    }
    const std::chrono::duration<double, std::milli> runtime = std::chrono::high_resolution_clock::now() - start;
    info("Problem %d runtime = %g ms", number, runtime.count());
}

int selftest(const std::string& dir);

} // namespace

int main(int argc, char** argv) {
    if (argc >= 3 && std::string(argv[1]) == "--selftest") return selftest(argv[2]);
    g_logfile = std::fopen("ps2.log", "a");
    const std::string config_path = argc >= 2 ? argv[1] : "../config/ps2.yaml";
    try {
        Config cfg = load_config(config_path);
        info("Loaded runtime configuration from \"%s\"", config_path.c_str());
        if (!cfg.use_gpu) {
            error("use_gpu_disparity is false: this build has no CPU disparity path (the reference's serial:: code is its test oracle)");
            return -1;
        }
        if (cfg.device != 0) setenv("STEREO_B200_DEVICE", std::to_string(cfg.device).c_str(), 1);
        std::unique_ptr<sb::MultiGpu> mgpu;
        if (cfg.gpus > 1) {
            std::vector<int> devs;
            for (int k = 0; k < cfg.gpus; ++k) devs.push_back(cfg.device + k);
            mgpu.reset(new sb::MultiGpu(devs));
            g_mgpu = mgpu.get();
            info("Sharding every pair over %d GPUs in row bands", mgpu->deviceCount());
        } else {
            sb::default_ctx();                                   // context creation = common::warmup() (main.cpp:346-349)
        }
        file_info("GPU warmup done");
        ps2::CvRng rng(cfg.noise_seed);
        const auto start = std::chrono::high_resolution_clock::now();
        using V = Variant;
        run_problem(cfg, rng, 1, "pair0", false, STEREO_COST_SSD, "disparitySSDKernel", {{'a', V::Clean}}, false);
        run_problem(cfg, rng, 2, "pair1", true, STEREO_COST_SSD, "disparitySSDKernel", {{'a', V::Clean}}, true);
        run_problem(cfg, rng, 3, "pair1", true, STEREO_COST_SSD, "disparitySSDKernel", {{'a', V::Noisy}, {'b', V::Contrast}}, true);
        run_problem(cfg, rng, 4, "pair1", true, STEREO_COST_NCORR, "disparityNCorrKernel", {{'a', V::Clean}, {'b', V::Noisy}, {'c', V::Contrast}}, true);
        run_problem(cfg, rng, 5, "pair2", true, STEREO_COST_NCORR, "disparityNCorrKernel", {{'a', V::Clean}}, true);
        const std::chrono::duration<double, std::milli> runtime = std::chrono::high_resolution_clock::now() - start;
        info("Total runtime: %g ms", runtime.count());
        g_mgpu = nullptr;
        g_gray.clear();
    } catch (const std::exception& e) {
        error("%s", e.what());
        error("Configuration load failed!");
        return -1;                                               // the reference exit(-1)s (lib/Config.cpp:39-49)
    }
    return 0;
}

namespace {
// Host-side self test: exercises the YAML subset, PNG round trip and every preprocessing step on files
// in `dir` prepared by the test (in.yaml, gray.png, colour.png, disp.bin) and writes its results next
// to them for comparison with executed OpenCV.
int selftest(const std::string& dir) {
    try {
        ps2::YamlDoc y = ps2::YamlDoc::load(dir + "/in.yaml");
        FILE* f = std::fopen((dir + "/yaml.txt").c_str(), "w");
        for (const auto& kv : y.values) std::fprintf(f, "%s=%s\n", kv.first.c_str(), kv.second.c_str());
        std::fclose(f);
        auto dump = [&](const std::string& name, const void* p, size_t n) { FILE* o = std::fopen((dir + "/" + name).c_str(), "wb"); std::fwrite(p, 1, n, o); std::fclose(o); };
        const ps2::Image8 gray = ps2::imread(dir + "/gray.png"), colour = ps2::imread(dir + "/colour.png");
        std::printf("gray %d %d %d\ncolour %d %d %d\n", gray.rows, gray.cols, gray.channels, colour.rows, colour.cols, colour.channels);
        dump("gray.raw", gray.data.data(), gray.data.size());
        dump("colour.raw", colour.data.data(), colour.data.size());
        ps2::imwrite_gray(dir + "/gray_out.png", gray.data.data(), gray.rows, gray.cols, size_t(gray.cols));
        const sb::Mat g14 = ps2::rgb2gray_on_bgr_as_float(colour, 14);
        dump("rgb2gray14.f32", g14.data, size_t(g14.rows) * g14.step);
        const sb::Mat g = ps2::rgb2gray_on_bgr_as_float(colour, 15);
        dump("rgb2gray.f32", g.data, size_t(g.rows) * g.step);
        ps2::CvRng rng(0xffffffffu);
        sb::Mat n1, n2;
        ps2::add_noise(rng, g, g, 0.f, 10.f, n1, n2);
        dump("noisy1.f32", n1.data, size_t(n1.rows) * n1.step);
        dump("noisy2.f32", n2.data, size_t(n2.rows) * n2.step);
        const sb::Mat c = ps2::scaled(g, 1.1f);
        dump("contrast.f32", c.data, size_t(c.rows) * c.step);
        // disp.bin: int32 rows, cols, then rows*cols int8
        FILE* d = std::fopen((dir + "/disp.bin").c_str(), "rb");
        int32_t hw[2];
        if (!d || std::fread(hw, 4, 2, d) != 2) throw std::runtime_error("disp.bin missing");
        sb::Mat disp(hw[0], hw[1], sb::S8C1);
        if (std::fread(disp.data, 1, size_t(hw[0]) * hw[1], d) != size_t(hw[0]) * hw[1]) throw std::runtime_error("disp.bin truncated");
        std::fclose(d);
        const sb::Mat n8 = ps2::normalize_minmax_u8(disp);
        dump("norm.u8", n8.data, size_t(n8.rows) * n8.step);
        const sb::Mat inv = ps2::inverted(n8);
        dump("inv.u8", inv.data, size_t(inv.rows) * inv.step);
        std::printf("selftest ok\n");
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "selftest failed: %s\n", e.what());
        return 1;
    }
}
} // namespace
