// orbx_api.cu — the C ABI (include/orbx.h): handle, geometry, arenas, pipeline orchestration.
// Host-side restatements cited inline: ORBextractor ctor (reference ORBextractor.cpp:409-469),
// ComputePyramid sizes (:1173-1174), the cell grid (:789-803), quadtree roots (:559-560) and the
// coefficient tables cv::resize builds for INTER_LINEAR.
#include "orbx_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

static std::string g_create_err;
static orbx_status grow(orbx_handle *h, uint8_t **p, size_t *cap, size_t need);
static const void *mapped_device_view(const void *host);
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool orbx_optin_smem(orbx_handle *h, const void *func, size_t smem)
{
    if (smem <= 48 * 1024) return true;
    if (smem > h->smem_optin_max) return false;
    static std::mutex mu;
    static std::vector<std::pair<std::pair<const void *, int>, size_t>> granted;      // (function, device) -> largest size set
    std::lock_guard<std::mutex> lock(mu);
    for (auto &g : granted) if (g.first.first == func && g.first.second == h->device) {
        if (smem <= g.second) return true;
        if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
        g.second = smem;
        return true;
    }
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
    granted.push_back({{func, h->device}, smem});
    return true;
}

extern "C" const char *orbx_version(void) { return "orbx 0.1 (sm_100a)"; }

extern "C" void orbx_default_params(orbx_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->nfeatures = 1000; p->scale_factor = 1.2f; p->nlevels = 8; p->ini_th_fast = 20; p->min_th_fast = 7;   // frontend.cpp:205-211
    p->depth_min = 0.3f; p->depth_max = 3.0f;                                                             // frontend.cpp:241-242
    p->max_width = 1280; p->max_height = 720; p->max_batch = 1; p->max_keypoints = 0; p->cand_divisor = 0; p->device = 0;
}

extern "C" const char *orbx_last_error(const orbx_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

// ---- ORBextractor ctor tables — ORBextractor.cpp:409-469 ----
static void build_tables(orbx_handle *h)
{
    const orbx_params &p = h->prm;
    const double scaleFactor = (double)p.scale_factor;
    h->scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < p.nlevels; i++) {
        h->scale[i] = (float)((double)h->scale[i - 1] * scaleFactor);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < p.nlevels; i++) { h->inv_scale[i] = 1.0f / h->scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    const float factor = (float)(1.0f / scaleFactor);
    float nDesired = p.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)p.nlevels));
    int sum = 0;
    for (int l = 0; l < p.nlevels - 1; l++) { h->nfeat[l] = cv_round_f(nDesired); sum += h->nfeat[l]; nDesired *= factor; }
    h->nfeat[p.nlevels - 1] = std::max(p.nfeatures - sum, 0);
    int v, v0;
    const int vmax = (int)floorf(ORBX_HALF_PATCH * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(ORBX_HALF_PATCH * sqrtf(2.f) / 2);
    const double hp2 = ORBX_HALF_PATCH * ORBX_HALF_PATCH;
    for (v = 0; v <= vmax; ++v) h->umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (v = ORBX_HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
        h->umax[v] = v0; ++v0;
    }
}

// cv::resize INTER_LINEAR coefficient tables
static void resize_table(int ssize, int dsize, bool horizontal, ResizeTab *out)
{
    const double inv_scale = (double)dsize / ssize, scale = 1. / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (horizontal) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        }
        const int a0 = cv_round_f((1.f - f) * 2048.f), a1 = cv_round_f(f * 2048.f);
        out[d].ofs = s;
        out[d].a0 = (int16_t)std::min(32767, std::max(-32768, a0));
        out[d].a1 = (int16_t)std::min(32767, std::max(-32768, a1));
    }
}

// geometry for one input size; returns false when the reference itself is undefined for it
static bool build_geometry(const orbx_handle *h, int w, int hgt, FrameGeom &G, std::vector<ResizeTab> *xt, std::vector<ResizeTab> *yt,
                           std::vector<uint32_t> *strip_tab = nullptr, std::vector<uint32_t> *blur_tab = nullptr,
                           std::vector<uint32_t> *cell_tab = nullptr)
{
    memset(&G, 0, sizeof(G));
    const orbx_params &p = h->prm;
    G.nlevels = p.nlevels; G.width = w; G.height = hgt;
    const int cdiv = p.cand_divisor > 0 ? p.cand_divisor : 16;
    size_t off = 0, boff = 0, coff = 0; int soff = 0, cells = 0, tiles = 0, ncmax = 8, strips = 0, max_hcell = 1, max_wcell = 1, cells_valid = 0;
    for (int l = 0; l < p.nlevels; l++) {
        LevelGeom &g = G.lv[l];
        g.w = cv_round_f((float)w * h->inv_scale[l]);                     // ORBextractor.cpp:1174
        g.h = cv_round_f((float)hgt * h->inv_scale[l]);
        if (g.w < 1 || g.h < 1) return false;
        g.pitch = (int)align_up((size_t)g.w, 128);
        if (l > 0) { g.off = off; off += align_up((size_t)g.pitch * g.h, 256); }
        g.bpitch = g.pitch; g.boff = boff; boff += align_up((size_t)g.bpitch * g.h, 256);
        const int W = g.w - 2 * ORBX_BORDER, H = g.h - 2 * ORBX_BORDER;    // maxBorder - minBorder, :786-795
        // Where the reference itself leaves defined behaviour (pinned against the compiled reference, tests/test_oracle_vs_ref.py,
        // same rule as oracle orc_geometry_status): a level with exactly 32 rows or with borders of opposite sign makes the root
        // count nIni (:559) infinite or negative and vpIniNodes.resize(nIni) throws (:566); nIni == 0 with a cell grid indexes an
        // empty vector (:586, checked below).  Such frame sizes are refused, never approximated.
        if (H == 0 || (int)roundf((float)W / (float)H) < 0) return false;
        g.scale = h->scale[l];
        g.size = (float)(int)(ORBX_PATCH * h->scale[l]);                  // :880
        g.N = h->nfeat[l];
        if (W >= ORBX_CELL_W && H >= ORBX_CELL_W) {
            const float width = (float)W, height = (float)H;
            g.ncols = (int)(width / (float)ORBX_CELL_W); g.nrows = (int)(height / (float)ORBX_CELL_W);   // :799-800
            g.wcell = (int)ceilf(width / g.ncols); g.hcell = (int)ceilf(height / g.nrows);               // :801-802
            g.nini = (int)roundf((float)W / H);                                                           // :559
            if (g.nini < 1 || g.nini > 64) return false;          // reference divides by zero for nini == 0
            g.hx = (float)W / g.nini;                                                                     // :560
        } else { g.ncols = g.nrows = 0; g.wcell = 1; g.hcell = std::max(8, std::min(32, g.h)); g.nini = 0; g.hx = 1.f; }   // no cell grid: hcell only sizes the blur tiles
        max_hcell = std::max(max_hcell, g.hcell);
        g.cell_first = cells; cells += g.ncols * g.nrows;
        if (g.ncols > 0) {
            const int cps_max = std::max(1, std::min(8, ORBX_FAST_MAX_W / g.wcell));
            g.strips_per_row = (g.ncols + cps_max - 1) / cps_max;
            g.cells_per_strip = (g.ncols + g.strips_per_row - 1) / g.strips_per_row;
            max_hcell = std::max(max_hcell, g.hcell);
        } else { g.strips_per_row = 0; g.cells_per_strip = 1; }
        g.strip_first = strips; strips += g.strips_per_row * g.nrows;
        // cell descriptors for k_fast_cells: level:4 | cell row:14 | cell column:14.  Cells the reference skips
        // (iniY >= maxBorderY-3, iniX >= maxBorderX-6; ORBextractor.cpp:811-816) are left out.
        for (int ci = 0; ci < g.nrows; ci++) for (int cj = 0; cj < g.ncols; cj++) {
            if (ORBX_BORDER + ci * g.hcell >= g.h - ORBX_BORDER - 3 || ORBX_BORDER + cj * g.wcell >= g.w - ORBX_BORDER - 6) continue;
            cells_valid++;
            if (cell_tab) cell_tab->push_back((uint32_t)l | ((uint32_t)ci << 4) | ((uint32_t)cj << 18));
        }
        if (g.ncols > 0) max_wcell = std::max(max_wcell, g.wcell);
        // strip descriptors (unused by the warp-per-cell FAST kernel, kept for the geometry report): level:4 | cells:4 | cell row:12 | first cell column:12.  Cells the reference
        // skips (iniY >= maxBorderY-3, iniX >= maxBorderX-6; ORBextractor.cpp:811-816) are trimmed here.
        if (strip_tab) for (int ci = 0; ci < g.nrows; ci++) for (int sj = 0; sj < g.strips_per_row; sj++) {
            const int maxBX = g.w - ORBX_BORDER, maxBY = g.h - ORBX_BORDER;
            const int cj0 = sj * g.cells_per_strip;
            int nc = std::min(g.cells_per_strip, g.ncols - cj0);
            if (ORBX_BORDER + ci * g.hcell >= maxBY - 3) continue;
            while (nc > 0 && ORBX_BORDER + (cj0 + nc - 1) * g.wcell >= maxBX - 6) nc--;
            if (nc <= 0) continue;
            strip_tab->push_back((uint32_t)l | ((uint32_t)nc << 4) | ((uint32_t)ci << 8) | ((uint32_t)cj0 << 20));
        }
        // blur tiles: 256 columns x hCell rows (the box of the level's tensor map is hCell + 6 rows); level:4 | column:12 | row:16
        g.blur_tx = (g.w + ORBX_BLUR_TW - 1) / ORBX_BLUR_TW; g.blur_ty = (g.h + g.hcell - 1) / g.hcell;
        g.blur_first = tiles; tiles += g.blur_tx * g.blur_ty;
        if (blur_tab) for (int ty = 0; ty < g.blur_ty; ty++) for (int tx = 0; tx < g.blur_tx; tx++)
            blur_tab->push_back((uint32_t)l | ((uint32_t)tx << 4) | ((uint32_t)ty << 16));
        g.cand_cap = (int)align_up((size_t)std::max(4096, g.w * g.h / cdiv), 64);
        g.cand_off = coff; coff += (size_t)g.cand_cap;
        g.sel_cap = (int)align_up((size_t)std::max(g.N + 4, 4 * g.nini + 4), 8);
        g.sel_off = soff; soff += g.sel_cap;
        ncmax = std::max(ncmax, g.sel_cap);
        if (l > 0 && xt && yt) {
            g.xtab_off = (int)xt->size(); xt->resize(xt->size() + g.w);
            resize_table(G.lv[l - 1].w, g.w, true, xt->data() + g.xtab_off);
            g.ytab_off = (int)yt->size(); yt->resize(yt->size() + g.h);
            resize_table(G.lv[l - 1].h, g.h, false, yt->data() + g.ytab_off);
        }
    }
    // output tile of the resize kernel (k_pyramid.cu): the largest (multiple of 16) x (multiple of 8) tile, at most 192 x 64,
    // whose source window fits the 288-byte x ORBX_RZ_BOX_ROWS TMA box at every level
    int rtw = 192, rth = 64;
    if (xt && yt) for (bool fits = false; !fits;) {
        fits = true;
        for (int l = 1; l < p.nlevels && fits; l++) {
            const LevelGeom &g = G.lv[l];
            const ResizeTab *tx = xt->data() + g.xtab_off, *ty = yt->data() + g.ytab_off;
            for (int x0 = 0; x0 < g.w && fits; x0 += rtw) {
                const int x1 = std::min(x0 + rtw, g.w) - 1;
                if (tx[x1].ofs + 1 - (tx[x0].ofs & ~15) >= 288) fits = false;
            }
            for (int y0 = 0; y0 < g.h && fits; y0 += rth) {
                const int y1 = std::min(y0 + rth, g.h) - 1;
                const int lo = std::max(0, std::min(ty[y0].ofs, G.lv[l - 1].h - 1)), hi = std::max(0, std::min(ty[y1].ofs + 1, G.lv[l - 1].h - 1));
                if (hi - lo + 1 > 80) fits = false;
            }
        }
        if (!fits) { if (rtw > 16) rtw -= 16; if (rth > 8) rth -= 8; if (rtw <= 16 && rth <= 8) return false; }
    }
    G.rz_tw = rtw; G.rz_th = rth;
    // batch kernel (k_resize_linear2): 128 x 64 output tiles, source window in a 192-byte x 80-row box
    G.rz2_ok = (xt && yt) ? 1 : 0;
    if (xt && yt) for (int l = 1; l < p.nlevels && G.rz2_ok; l++) {
        const LevelGeom &g = G.lv[l];
        const ResizeTab *tx = xt->data() + g.xtab_off, *ty = yt->data() + g.ytab_off;
        for (int x0 = 0; x0 < g.w && G.rz2_ok; x0 += 128) if (tx[std::min(x0 + 128, g.w) - 1].ofs + 1 - (tx[x0].ofs & ~15) >= 192 - 8) G.rz2_ok = 0;   // - 8: the third word rz_hrow reads
        for (int y0 = 0; y0 < g.h && G.rz2_ok; y0 += 64) {
            const int y1 = std::min(y0 + 64, g.h) - 1;
            const int lo = std::max(0, std::min(ty[y0].ofs, G.lv[l - 1].h - 1)), hi = std::max(0, std::min(ty[y1].ofs + 1, G.lv[l - 1].h - 1));
            if (hi - lo + 1 > 80) G.rz2_ok = 0;
        }
    }
    G.total_cells = cells; G.total_blur_tiles = tiles; G.total_strips = std::max(strips, 1); G.max_hcell = max_hcell; G.max_wcell = max_wcell; G.total_cells_valid = cells_valid;
    G.pyr_bytes = std::max<size_t>(off, 256); G.blur_bytes = boff; G.cand_entries = coff; G.sel_entries = soff; G.node_cap_max = ncmax;
    return true;
}

static orbx_status set_geometry(orbx_handle *h, int w, int hgt)
{
    if (h->geo.width == w && h->geo.height == hgt) return ORBX_OK;
    if (w > h->prm.max_width || hgt > h->prm.max_height) { h->err = "frame larger than max_width x max_height"; return ORBX_E_INVALID; }
    if (w > 4096 + 2 * ORBX_BORDER || hgt > 4096 + 2 * ORBX_BORDER) { h->err = "frame larger than 4128 px"; return ORBX_E_UNSUPPORTED; }
    FrameGeom G; std::vector<ResizeTab> xt, yt; std::vector<uint32_t> strips, btiles, ctab;
    if (!build_geometry(h, w, hgt, G, &xt, &yt, &strips, &btiles, &ctab)) { h->err = "unsupported frame geometry: the reference throws or faults here (a pyramid level vanishes, has exactly 32 rows, or its aspect ratio gives a negative or zero quadtree root count, ORBextractor.cpp:559-586), or it has more than 64 roots"; return ORBX_E_UNSUPPORTED; }
    // the quadtree keeps its node table in shared memory with 16-bit node ids (k_quadtree.cu): refuse here, with a clear message,
    // what would otherwise surface as a launch failure on the first extraction
    if (G.node_cap_max >= 65535 || orbx_quadtree_smem(G.node_cap_max) > h->smem_optin_max) {
        h->err = "nfeatures per level too large for the quadtree's shared-memory node table (" + std::to_string(orbx_quadtree_smem(G.node_cap_max)) +
                 " bytes needed, " + std::to_string(h->smem_optin_max) + " available per block)";
        return ORBX_E_UNSUPPORTED;
    }
    const size_t B = (size_t)h->prm.max_batch;
    if (G.pyr_bytes * B > h->pyr_cap || G.blur_bytes * B > h->blur_cap || G.cand_entries * B > h->cand_cap ||
        (size_t)G.sel_entries * B > h->sel_cap || (int)xt.size() > h->tab_cap || (int)yt.size() > h->tab_cap ||
        G.sel_entries > h->max_kp || (int)strips.size() > h->strip_cap || (int)btiles.size() > h->blur_tile_cap || (int)ctab.size() > h->cell_cap) { h->err = "frame geometry does not fit the arenas sized at create"; return ORBX_E_INVALID; }
    G.total_strips = (int)strips.size();
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBX_CUDA(h, cudaMemcpy(h->d_geo, &G, sizeof(G), cudaMemcpyHostToDevice));
    if (!xt.empty()) ORBX_CUDA(h, cudaMemcpy(h->d_xtab, xt.data(), xt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice));
    if (!yt.empty()) ORBX_CUDA(h, cudaMemcpy(h->d_ytab, yt.data(), yt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice));
    if (!strips.empty()) ORBX_CUDA(h, cudaMemcpy(h->d_strips, strips.data(), strips.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (!ctab.empty()) {
        std::vector<uint4> recs;
        orbx_build_fast_cells(G, ctab, recs);
        ORBX_CUDA(h, cudaMemcpy(h->d_cells, recs.data(), recs.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    }
    if (!btiles.empty()) ORBX_CUDA(h, cudaMemcpy(h->d_blur_tiles, btiles.data(), btiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    h->dgeo.ntiles = 0;                                    // 0 tiles = this geometry runs the warp-per-cell kernel
    if (h->dense_ok) {
        DenseGeom D; std::vector<uint4> dt;
        orbx_build_fast_dense(G, h->prm.cand_divisor, D, &dt);
        if (D.ntiles > 0 && D.ntiles <= h->dtile_cap && D.map_bytes * B <= h->smap_cap && D.cl_entries * B <= h->clist_cap) {
            ORBX_CUDA(h, cudaMemcpy(h->d_dtiles, dt.data(), dt.size() * sizeof(uint4), cudaMemcpyHostToDevice));
            h->dgeo = D;
        }
    }
    h->geo_serial++;
    G.total_strips = (int)strips.size();
    h->geo = G; h->pyr_slab = G.pyr_bytes; h->blur_slab = G.blur_bytes; h->tmap_valid = false; h->alt.tmap_valid = false;
    return ORBX_OK;
}

extern "C" void orbx_destroy(orbx_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    orbx_cvorb_destroy(h);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    void *dev[] = { h->d_bgr, h->d_cells, h->d_blur_tiles, h->d_strips, h->d_prev_desc, h->d_prev_count, h->d_geo, h->d_xtab, h->d_ytab, h->d_pyr, h->d_blur, h->d_in, h->d_depth_in, h->d_cand, h->d_cand2, h->d_qtmp,
                    h->d_owner, h->d_owner2, h->d_ncand, h->d_sel, h->d_nsel, h->d_kps_all, h->d_desc_all, h->d_count_all,
                    h->d_kps_out, h->d_desc_out, h->d_count_out, h->d_boxes, h->d_box_off, h->d_status_base, h->d_mpart, h->d_mq, h->d_mt, h->d_mout, h->d_mcount,
                    h->d_dtiles, h->d_smap, h->d_clist, h->d_dense_zero, h->d_retry };
    for (void *p : dev) if (p) cudaFree(p);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->ev_a) cudaEventDestroy(h->ev_a);
    if (h->ev_b) cudaEventDestroy(h->ev_b);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_fork0) cudaEventDestroy(h->ev_fork0);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_lane_fork) cudaEventDestroy(h->ev_lane_fork);
    if (h->ev_lane_join) cudaEventDestroy(h->ev_lane_join);
    if (h->ev_prev_ready) cudaEventDestroy(h->ev_prev_ready);
    if (h->alt.allocated) {
        void *lane[] = { h->alt.d_pyr, h->alt.d_cand, h->alt.d_cand2, h->alt.d_qtmp, h->alt.d_owner, h->alt.d_owner2, h->alt.d_ncand, h->alt.d_nsel, h->alt.d_sel,
                         h->alt.d_kps_all, h->alt.d_desc_all, h->alt.d_count_all, h->alt.d_mpart };
        for (void *p : lane) if (p) cudaFree(p);
        if (h->alt.stream) cudaStreamDestroy(h->alt.stream);
    }
    for (int i = 0; i < 2; i++) { if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]); if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]); if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]); }
    delete h;
}

#define CREATE_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    g_create_err = std::string(#call) + ": " + cudaGetErrorString(e_); orbx_destroy(h); return ORBX_E_CUDA; } } while (0)

extern "C" orbx_status orbx_create(const orbx_params *pp, orbx_handle **out)
{
    if (!pp || !out) { g_create_err = "null argument"; return ORBX_E_INVALID; }
    *out = nullptr;
    orbx_params p = *pp;
    if (p.nlevels < 1 || p.nlevels > ORBX_MAX_LEVELS || p.nfeatures < 1 || !(p.scale_factor > 1.0f) ||
        p.max_width < 1 || p.max_height < 1 || p.max_batch < 1 || p.ini_th_fast < 0 || p.min_th_fast < 0 ||
        p.ini_th_fast > 255 || p.min_th_fast > 255) { g_create_err = "invalid orbx_params"; return ORBX_E_INVALID; }
    if (p.profile != ORBX_PROFILE_SLAM && p.profile != ORBX_PROFILE_CVORB) { g_create_err = "unknown orbx_params.profile"; return ORBX_E_INVALID; }
    if (p.scale_factor == 2.0f) { g_create_err = "scale_factor 2.0 takes cv::resize's INTER_AREA fast path, not implemented"; return ORBX_E_UNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { g_create_err = "no CUDA device: this path has no CPU fallback"; return ORBX_E_CUDA; }
    if (p.device < 0 || p.device >= ndev) { g_create_err = "device ordinal out of range"; return ORBX_E_INVALID; }
    orbx_handle *h = new orbx_handle();
    h->prm = p; h->device = p.device; h->launches = 0; h->geo.width = -1; h->geo.height = -1;
    h->prof_on = 0; h->prof_n = 0; h->prev_valid = 0; h->opt_fused_blur = 1; h->opt_filter_first = 1; h->blur_valid = false; h->opt_pdl = 1; h->opt_overlap = 0; h->opt_match_mma = 1; h->in_overlap = false; h->ev_after_pyramid = nullptr;
    memset(&h->alt, 0, sizeof(h->alt));
    memset(h->prof_ms, 0, sizeof(h->prof_ms)); memset(h->prof_cnt, 0, sizeof(h->prof_cnt));
    CREATE_CUDA(cudaSetDevice(p.device));
    cudaDeviceProp prop;
    CREATE_CUDA(cudaGetDeviceProperties(&prop, p.device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin_max = prop.sharedMemPerBlockOptin;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);          // (least, greatest): the dependent chain outranks the filler stream
    CREATE_CUDA(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));
    CREATE_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, prio_lo));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_fork0, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_lane_fork, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_lane_join, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_prev_ready, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
        CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    build_tables(h);
    upload_umax(h->umax);
    FrameGeom G;
    if (!build_geometry(h, p.max_width, p.max_height, G, nullptr, nullptr)) {
        g_create_err = "unsupported max_width x max_height geometry"; orbx_destroy(h); return ORBX_E_UNSUPPORTED;
    }
    const size_t B = (size_t)p.max_batch;
    // host-buffer calls are pipelined in chunks of `chunk` frames over two input and two output slots
    h->chunk = p.host_chunk > 0 ? std::min(p.host_chunk, p.max_batch) : std::max(1, std::min(32, p.max_batch / 4));
    h->seq = 0; h->pending[0].active = h->pending[1].active = false;
    const size_t S = 2 * B;                                       // the staging arenas hold two slots of max_batch frames
    // arenas are sized for the max geometry with 12% headroom so that smaller frames with unlucky padding still fit
    h->pyr_cap = (G.pyr_bytes + G.pyr_bytes / 8 + 4096) * B; h->blur_cap = (G.blur_bytes + G.blur_bytes / 8 + 4096) * B;
    h->cand_cap = (G.cand_entries + G.cand_entries / 8 + 4096 * p.nlevels) * B;
    h->sel_cap = (size_t)(G.sel_entries + 64 * p.nlevels) * B;
    h->max_kp = std::max(p.max_keypoints, G.sel_entries + 64 * p.nlevels);
    h->tab_cap = 2 * (p.max_width + p.max_height) * 6 + 1024;
    h->in_cap = align_up((size_t)p.max_width, 128) * p.max_height * S;
    h->depth_cap = align_up((size_t)p.max_width * 2, 128) * p.max_height * S;
    CREATE_CUDA(cudaMalloc(&h->d_geo, sizeof(FrameGeom)));
    h->strip_cap = G.total_cells + 64 * p.nlevels;
    CREATE_CUDA(cudaMalloc(&h->d_strips, sizeof(uint32_t) * h->strip_cap));
    h->cell_cap = G.total_cells + G.total_cells / 4 + 64 * p.nlevels;
    CREATE_CUDA(cudaMalloc(&h->d_cells, 2 * sizeof(uint4) * h->cell_cap));
    h->opt_fast_dense = 0; h->opt_dense_fuse = 0; h->dense_ok = false;          // its arenas are allocated when ORBX_OPT_FAST_DENSE is first switched on (dense_alloc)
    h->blur_tile_cap = G.total_blur_tiles + G.total_blur_tiles / 4 + 64 * p.nlevels;
    CREATE_CUDA(cudaMalloc(&h->d_blur_tiles, sizeof(uint32_t) * h->blur_tile_cap));
    CREATE_CUDA(cudaMalloc(&h->d_xtab, sizeof(ResizeTab) * h->tab_cap));
    CREATE_CUDA(cudaMalloc(&h->d_ytab, sizeof(ResizeTab) * h->tab_cap));
    CREATE_CUDA(cudaMalloc(&h->d_pyr, h->pyr_cap));
    CREATE_CUDA(cudaMalloc(&h->d_blur, h->blur_cap));
    CREATE_CUDA(cudaMalloc(&h->d_in, h->in_cap));
    CREATE_CUDA(cudaMalloc(&h->d_depth_in, h->depth_cap));
    CREATE_CUDA(cudaMalloc(&h->d_cand, h->cand_cap * sizeof(uint32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_cand2, h->cand_cap * sizeof(uint32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_qtmp, h->cand_cap * sizeof(uint32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_owner, h->cand_cap * sizeof(uint16_t)));
    CREATE_CUDA(cudaMalloc(&h->d_owner2, h->cand_cap * sizeof(uint16_t)));
    CREATE_CUDA(cudaMalloc(&h->d_ncand, (B * ORBX_MAX_LEVELS + 4) * sizeof(int32_t)));   // + the FAST work counter
    CREATE_CUDA(cudaMalloc(&h->d_nsel, B * ORBX_MAX_LEVELS * sizeof(int32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_sel, h->sel_cap * sizeof(uint32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_kps_all, B * h->max_kp * sizeof(orbx_keypoint)));
    CREATE_CUDA(cudaMalloc(&h->d_desc_all, B * h->max_kp * ORBX_DESC_BYTES));
    CREATE_CUDA(cudaMalloc(&h->d_count_all, B * sizeof(int32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_kps_out, S * h->max_kp * sizeof(orbx_keypoint)));
    CREATE_CUDA(cudaMalloc(&h->d_desc_out, S * h->max_kp * ORBX_DESC_BYTES));
    CREATE_CUDA(cudaMalloc(&h->d_count_out, S * sizeof(int32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_prev_desc, (size_t)h->max_kp * ORBX_DESC_BYTES));
    CREATE_CUDA(cudaMalloc(&h->d_prev_count, sizeof(int32_t)));
    CREATE_CUDA(cudaMemset(h->d_prev_count, 0, sizeof(int32_t)));
    h->boxes_cap = std::max<size_t>(256, 16 * B);
    CREATE_CUDA(cudaMalloc(&h->d_boxes, 2 * (size_t)h->boxes_cap * sizeof(orbx_box)));
    CREATE_CUDA(cudaMalloc(&h->d_box_off, 2 * (B + 1) * sizeof(int32_t)));
    CREATE_CUDA(cudaMalloc(&h->d_status_base, 4 * sizeof(int32_t)));
    h->d_status = h->d_status_base;
    CREATE_CUDA(cudaMalloc(&h->d_mcount, sizeof(int32_t) * 4));
    CREATE_CUDA(cudaMemset(h->d_status_base, 0, 4 * sizeof(int32_t)));
    CREATE_CUDA(cudaMemset(h->d_nsel, 0, B * ORBX_MAX_LEVELS * sizeof(int32_t)));
    h->h_out_bytes = B * ((size_t)h->max_kp * (sizeof(orbx_keypoint) + ORBX_DESC_BYTES) + 64);
    CREATE_CUDA(cudaMallocHost(&h->h_out, h->h_out_bytes));
    CREATE_CUDA(cudaMallocHost(&h->h_status, 64));
    memset(h->h_status, 0, 64);
    orbx_status st = set_geometry(h, p.max_width, p.max_height);
    if (st != ORBX_OK) { g_create_err = h->err; orbx_destroy(h); return st; }
    if (p.profile == ORBX_PROFILE_CVORB) {
        if ((st = orbx_cvorb_create(h)) != ORBX_OK) { g_create_err = h->err; orbx_destroy(h); return st; }
        // cv::ORB's per-level scale is pow(scaleFactor, level) in one step, not the extractor's running product (orb.cpp getScale)
        for (int l = 0; l < p.nlevels; l++) {
            h->scale[l] = (float)pow((double)p.scale_factor, (double)l);
            h->inv_scale[l] = 1.0f / h->scale[l]; h->sigma2[l] = h->scale[l] * h->scale[l]; h->inv_sigma2[l] = 1.0f / h->sigma2[l];
        }
    }
    *out = h;
    return ORBX_OK;
}

// surfaces device-side error flags (capacity overflows) after a synchronisation point
static orbx_status check_device_status(orbx_handle *h)
{
    ORBX_CUDA(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    const int s = h->h_status[0];
    if (s == 0) return ORBX_OK;
    ORBX_CUDA(h, cudaMemsetAsync(h->d_status, 0, sizeof(int32_t), h->stream));
    if (s & ORBX_DS_BAD_INDEX) { h->err = "a match query index lies outside the keypoint array"; return ORBX_E_INVALID; }
    if (s & ORBX_DS_INTERNAL) { h->err = "internal: a frame of the dense FAST kernel did not complete within its wait bound; results of this call are incomplete"; return ORBX_E_CUDA; }
    h->err = std::string("device capacity exceeded:") + ((s & ORBX_DS_CAND_OVERFLOW) ? " candidate list (lower cand_divisor)" : "") +
             ((s & ORBX_DS_NODE_OVERFLOW) ? " quadtree nodes / retained-corner list" : "") + ((s & ORBX_DS_KP_OVERFLOW) ? " keypoint output (raise cap / max_keypoints)" : "");
    return ORBX_E_CAPACITY;
}

extern "C" orbx_status orbx_sync(orbx_handle *h)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    return check_device_status(h);
}
// dense FAST formulation (k_fast_dense.cu): score maps + corner lists for max_batch frames of the largest geometry, allocated when the
// option is first switched on (0.8 GB at 128 frames of 1280 x 720)
static orbx_status dense_alloc(orbx_handle *h)
{
    if (h->dense_ok) return ORBX_OK;
    const orbx_params &p = h->prm;
    if (p.profile != ORBX_PROFILE_SLAM) { h->err = "ORBX_OPT_FAST_DENSE is a profile-S option"; return ORBX_E_UNSUPPORTED; }
    cudaSetDevice(h->device);
    FrameGeom G;
    if (!build_geometry(h, p.max_width, p.max_height, G, nullptr, nullptr)) { h->err = "max_width x max_height is not a supported geometry"; return ORBX_E_UNSUPPORTED; }
    const size_t B = (size_t)p.max_batch;
    DenseGeom D;
    orbx_build_fast_dense(G, p.cand_divisor, D, nullptr);
    h->dtile_cap = D.ntiles + D.ntiles / 4 + 64 * p.nlevels;
    h->smap_cap = (D.map_bytes + D.map_bytes / 8 + 65536) * B;
    h->clist_cap = (D.cl_entries + D.cl_entries / 8 + 4096 * p.nlevels) * B;
    h->dense_zero_bytes = (4 + B * ORBX_MAX_LEVELS + B) * sizeof(int32_t) + B * (size_t)h->cell_cap;
    if (cudaMalloc(&h->d_dtiles, sizeof(uint4) * h->dtile_cap) != cudaSuccess || cudaMalloc(&h->d_smap, h->smap_cap) != cudaSuccess ||
        cudaMalloc(&h->d_clist, h->clist_cap * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&h->d_dense_zero, h->dense_zero_bytes) != cudaSuccess ||
        cudaMalloc(&h->d_retry, B * (size_t)h->cell_cap * sizeof(int32_t)) != cudaSuccess) {
        cudaGetLastError();
        void *dn[] = { h->d_dtiles, h->d_smap, h->d_clist, h->d_dense_zero, h->d_retry };
        for (void *q : dn) if (q) cudaFree(q);
        h->d_dtiles = nullptr; h->d_smap = nullptr; h->d_clist = nullptr; h->d_dense_zero = nullptr; h->d_retry = nullptr;
        h->err = "out of device memory for the dense FAST arenas";
        return ORBX_E_CUDA;
    }
    h->dense_ok = true;
    h->geo.width = -1; h->geo.height = -1;                 // the next call rebuilds the geometry, now with the tile table
    return ORBX_OK;
}

extern "C" orbx_status orbx_set_option(orbx_handle *h, int32_t option, int32_t value)
{
    if (!h) return ORBX_E_INVALID;
    if (option == ORBX_OPT_SERIAL) { h->opt_serial = value ? 1 : 0; return ORBX_OK; }
    if (option == ORBX_OPT_FAST_CTAS) { h->opt_fast_ctas = value > 0 ? value : 0; return ORBX_OK; }
    if (option == ORBX_OPT_FUSED_BLUR) { h->opt_fused_blur = value ? 1 : 0; return ORBX_OK; }
    if (option == ORBX_OPT_FILTER_FIRST) { h->opt_filter_first = value ? 1 : 0; return ORBX_OK; }
    if (option == ORBX_OPT_PDL) { h->opt_pdl = value ? 1 : 0; return ORBX_OK; }
    if (option == ORBX_OPT_OVERLAP) { h->opt_overlap = value ? 1 : 0; return ORBX_OK; }
    if (option == ORBX_OPT_FAST_DENSE) {
        const int v = value < 0 ? 0 : (value > 3 ? 3 : value);
        if (v > 0) { const orbx_status st = dense_alloc(h); if (st != ORBX_OK) return st; }
        h->opt_fast_dense = v == 3 ? 2 : v;
        h->opt_dense_fuse = v == 3 ? 1 : 0;                      // 3: every call, the NMS as work items of the tile kernel (measured slower: k_fast_dense.cu)
        return ORBX_OK;
    }
    if (option == ORBX_OPT_MATCH_MMA) { h->opt_match_mma = value < 0 ? 0 : (value > 3 ? 3 : value); return ORBX_OK; }
    h->err = "unknown option"; return ORBX_E_INVALID;
}
extern "C" void *orbx_stream(orbx_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int64_t orbx_launch_count(const orbx_handle *h) { return h ? h->launches : 0; }

extern "C" int32_t orbx_get_levels(const orbx_handle *h) { return h ? h->prm.nlevels : 0; }
extern "C" float orbx_get_scale_factor(const orbx_handle *h) { return h ? (float)(double)h->prm.scale_factor : 0.f; }
extern "C" void orbx_get_scale_factors(const orbx_handle *h, float *o) { if (h && o) memcpy(o, h->scale, sizeof(float) * h->prm.nlevels); }
extern "C" void orbx_get_inverse_scale_factors(const orbx_handle *h, float *o) { if (h && o) memcpy(o, h->inv_scale, sizeof(float) * h->prm.nlevels); }
extern "C" void orbx_get_scale_sigma_squares(const orbx_handle *h, float *o) { if (h && o) memcpy(o, h->sigma2, sizeof(float) * h->prm.nlevels); }
extern "C" void orbx_get_inverse_scale_sigma_squares(const orbx_handle *h, float *o) { if (h && o) memcpy(o, h->inv_sigma2, sizeof(float) * h->prm.nlevels); }
extern "C" void orbx_get_features_per_level(const orbx_handle *h, int32_t *o) { if (h && o) for (int l = 0; l < h->prm.nlevels; l++) o[l] = h->nfeat[l]; }
extern "C" orbx_status orbx_level_size(const orbx_handle *h, int32_t w, int32_t hgt, int32_t level, int32_t *lw, int32_t *lh)
{
    if (!h || level < 0 || level >= h->prm.nlevels || !lw || !lh) return ORBX_E_INVALID;
    if (h->cv) { int a, b; orbx_cvorb_level_size(h, w, hgt, level, &a, &b); *lw = a; *lh = b; return ORBX_OK; }
    *lw = cv_round_f((float)w * h->inv_scale[level]); *lh = cv_round_f((float)hgt * h->inv_scale[level]);
    return ORBX_OK;
}

// ---- the extraction pipeline: ORBextractor::operator() (ORBextractor.cpp:1086-1167) on nframes frames ----
// YOLO boxes on the device: one list for the whole call (off == nullptr) or per-frame ranges [off[f], off[f+1]) of `boxes`
// (`base` is subtracted from the offsets: a pipeline chunk stages only its own boxes)
struct DevBoxes { const orbx_box *boxes; const int32_t *off; int base; int n; uint64_t drop_mask; };
static const DevBoxes kNoBoxes = { nullptr, nullptr, 0, 0, 0 };
static orbx_status pipeline_device(orbx_handle *h, bool track, const uint8_t *d_gray, int nframes, int width, int height, size_t step, size_t fstride,
                                   const uint16_t *d_depth, size_t dstep, size_t dfstride, const DevBoxes &BX,
                                   orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts,
                                   orbx_dmatch *d_matches, int32_t *d_mcounts, float max_dist);

static orbx_status run_pipeline(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride,
                                const uint16_t *d_depth, size_t dstep, size_t dfstride, const DevBoxes &BX,
                                orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts)
{
    const orbx_box *d_boxes = BX.boxes; const int nboxes = BX.n; const uint64_t drop_mask = BX.drop_mask;
    if (h->cv) {                                                  // profile C (cv::ORB): one frame at a time through orbx_cvorb.cu
        const bool filt = d_depth != nullptr || nboxes > 0;
        for (int f = 0; f < nframes; f++) {
            const orbx_status st = filt
                ? orbx_cvorb_run(h, l0 + (size_t)f * l0_fstride, h->geo.width, h->geo.height, l0_step, h->d_kps_all + (size_t)f * h->max_kp,
                                 h->d_desc_all + (size_t)f * h->max_kp * ORBX_DESC_BYTES, h->max_kp, h->d_count_all + f)
                : orbx_cvorb_run(h, l0 + (size_t)f * l0_fstride, h->geo.width, h->geo.height, l0_step, d_kps + (size_t)f * cap,
                                 d_desc + (size_t)f * cap * ORBX_DESC_BYTES, cap, d_counts + f);
            if (st != ORBX_OK) return st;
        }
        if (filt) launch_filter(h, nframes, d_depth, dstep, dfstride, d_boxes, BX.off, BX.base, nboxes, drop_mask, d_kps, d_desc, cap, d_counts);
        h->last_batch = nframes; h->last_l0 = l0; h->last_l0_step = l0_step; h->last_l0_fstride = l0_fstride;
        ORBX_CUDA(h, cudaGetLastError());
        return ORBX_OK;
    }
    const int nl = h->geo.nlevels;
    if (h->geo.total_cells_valid == 0) {
        // no level holds a FAST cell (every level is narrower than 67 px): the reference's cell loops do not execute and it returns no
        // keypoints (ORBextractor.cpp:799-806); nothing is launched, the pyramid is not built
        ORBX_CUDA(h, cudaMemsetAsync(d_counts, 0, (size_t)nframes * sizeof(int32_t), h->stream));
        h->last_batch = 0;
        return ORBX_OK;
    }
    ORBX_CUDA(h, cudaMemsetAsync(h->d_ncand, 0, ((size_t)h->prm.max_batch * ORBX_MAX_LEVELS + 4) * sizeof(int32_t), h->stream));   // corner counts + FAST work counter
    // Schedule.  The main stream carries the dependent chain pyramid -> FAST -> quadtree -> describe; the blur (needed only by
    // describe) runs on the low-priority aux stream as filler: level 0 depends on the input alone and starts beside the
    // pyramid's seven shrinking launches, the other levels start once the pyramid exists and fill what FAST and the
    // (latency-bound, low-occupancy) quadtree leave free.
    h->pdl_chain = nframes <= 8;
    const int l0_tiles = h->geo.lv[0].blur_tx * h->geo.lv[0].blur_ty;
    const bool side = !h->opt_serial && !h->opt_fused_blur;                       // a blur kernel on the aux stream
    if (side) {
        ORBX_CUDA(h, cudaEventRecord(h->ev_fork0, h->stream));
        ORBX_CUDA(h, cudaStreamWaitEvent(h->aux_stream, h->ev_fork0, 0));
        if (launch_blur(h, nframes, l0, l0_step, l0_fstride, h->aux_stream, 0, l0_tiles) != 0) { h->err = "cuTensorMapEncodeTiled failed (frame base/step must be 16-byte aligned)"; return ORBX_E_CUDA; }
    }
    if (launch_pyramid(h, nframes, l0, l0_step, l0_fstride) != 0) { h->err = "cuTensorMapEncodeTiled failed (frame base/step must be 16-byte aligned)"; return ORBX_E_CUDA; }   // ComputePyramid
    if (h->ev_after_pyramid) ORBX_CUDA(h, cudaEventRecord(h->ev_after_pyramid, h->stream));
    if (side) {
        ORBX_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
        ORBX_CUDA(h, cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
    }
    if (launch_fast(h, nframes, l0, l0_step, l0_fstride) != 0) { h->err = "cuTensorMapEncodeTiled failed (frame base/step must be 16-byte aligned)"; return ORBX_E_CUDA; }   // cell FAST
    if (!h->opt_fused_blur) {                                                      // GaussianBlur per level as its own kernel
        if (h->opt_serial) { if (launch_blur(h, nframes, l0, l0_step, l0_fstride, h->stream) != 0) { h->err = "cuTensorMapEncodeTiled failed"; return ORBX_E_CUDA; } }
        else if (launch_blur(h, nframes, l0, l0_step, l0_fstride, h->aux_stream, l0_tiles, -1) != 0) { h->err = "cuTensorMapEncodeTiled failed"; return ORBX_E_CUDA; }
    }
    h->blur_valid = !h->opt_fused_blur;
    if (side) ORBX_CUDA(h, cudaEventRecord(h->ev_join, h->aux_stream));
    if (launch_quadtree(h, nframes) != 0) { h->err = "quadtree node table exceeds the shared-memory opt-in limit"; return ORBX_E_UNSUPPORTED; }   // DistributeOctTree
    if (side) ORBX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    const bool filtered = d_depth != nullptr || nboxes > 0;
    if (!filtered) launch_describe_to(h, nframes, l0, l0_step, l0_fstride, d_kps, d_desc, cap, d_counts);
    else if (h->opt_fused_blur && h->opt_filter_first) {
        // Filter first.  The depth / box filter (frontend.cpp:503-527, backend.cpp:1011-1029) looks at a keypoint's POSITION only, which is known
        // once the quadtree has selected it; the reference computes the angle and the descriptor of every selected keypoint and then drops
        // rows.  Here the filter runs on the selected positions and leaves an ordered index list (k_keep_list), and the descriptor kernel works
        // on the survivors alone, writing their final rows: same output, a fifth less descriptor work on the RGB-D stream, no compaction copy.
        // The index list lives in the (otherwise unused) unfiltered-keypoint arena, which every half-batch lane owns a copy of.
        int32_t *d_map = reinterpret_cast<int32_t *>(h->d_kps_all);
        launch_keep_list(h, nframes, d_depth, dstep, dfstride, d_boxes, BX.off, BX.base, nboxes, drop_mask, d_map, h->max_kp, d_counts, cap);
        launch_describe_to(h, nframes, l0, l0_step, l0_fstride, d_kps, d_desc, cap, d_counts, d_map, h->max_kp);
    } else {
        launch_describe_to(h, nframes, l0, l0_step, l0_fstride, h->d_kps_all, h->d_desc_all, h->max_kp, h->d_count_all);
        launch_filter(h, nframes, d_depth, dstep, dfstride, d_boxes, BX.off, BX.base, nboxes, drop_mask, d_kps, d_desc, cap, d_counts);
    }
    h->last_batch = nframes; h->last_l0 = l0; h->last_l0_step = l0_step; h->last_l0_fstride = l0_fstride;
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

extern "C" orbx_status orbx_extract_batch_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                                 int32_t width, int32_t height, size_t step, size_t frame_stride,
                                                 const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                                 orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap, int32_t *d_counts)
{
    return orbx_extract_batch_boxes_device(h, d_gray, nframes, width, height, step, frame_stride, d_depth, dstep, dframe_stride,
                                           nullptr, nullptr, 0, 0, d_kps, d_desc, cap, d_counts);
}

extern "C" orbx_status orbx_extract_batch_boxes_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                                       int32_t width, int32_t height, size_t step, size_t frame_stride,
                                                       const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                                       const orbx_box *d_boxes, const int32_t *d_box_offsets, int32_t nboxes_total, uint64_t drop_mask,
                                                       orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap, int32_t *d_counts)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (!d_gray || width <= 0 || height <= 0) { h->err = "empty image"; return ORBX_E_EMPTY; }
    if (nframes < 1 || nframes > h->prm.max_batch || !d_kps || !d_desc || !d_counts || cap < 1) { h->err = "bad batch arguments"; return ORBX_E_INVALID; }
    if (step < (size_t)width || (step & 15) || ((uintptr_t)d_gray & 15) || (frame_stride & 15)) { h->err = "device frames need 16-byte aligned base, step and frame stride"; return ORBX_E_INVALID; }
    if (d_depth && ((dstep & 1) || dstep < (size_t)width * 2)) { h->err = "bad depth step"; return ORBX_E_INVALID; }
    orbx_status st = set_geometry(h, width, height);
    if (st != ORBX_OK) return st;
    if (nboxes_total < 0 || (nboxes_total > 0 && (!d_boxes || !d_box_offsets))) { h->err = "per-frame boxes need the box array and nframes + 1 offsets"; return ORBX_E_INVALID; }
    if (nboxes_total > 0 && cap > h->max_kp) { h->err = "cap_per_frame larger than the handle's max_keypoints"; return ORBX_E_INVALID; }
    const DevBoxes BX = { d_boxes, d_box_offsets, 0, nboxes_total, drop_mask };
    return pipeline_device(h, false, d_gray, nframes, width, height, step, frame_stride, d_depth, dstep, dframe_stride, BX, d_kps, d_desc, cap, d_counts, nullptr, nullptr, 0.f);
}

static orbx_status extract_one(orbx_handle *h, bool is_bgr, const uint8_t *gray, int32_t width, int32_t height, size_t step,
                               const uint16_t *depth, size_t dstep, const orbx_box *boxes, int32_t nboxes, uint64_t drop_mask,
                               orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n_out) *n_out = 0;
    if (!gray || width <= 0 || height <= 0) { h->err = "empty image"; return ORBX_E_EMPTY; }        // ORBextractor.cpp:1090-1091
    if (!kps || !desc || !n_out || cap < 0 || step < (size_t)width * (is_bgr ? 3 : 1) || nboxes < 0 || (nboxes > 0 && !boxes)) { h->err = "bad arguments"; return ORBX_E_INVALID; }
    if (depth && (dstep < (size_t)width * 2 || (dstep & 1) || ((uintptr_t)depth & 1))) { h->err = "bad depth buffer: rows of at least 2*width bytes, even step, 2-byte aligned"; return ORBX_E_INVALID; }
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    orbx_status st = set_geometry(h, width, height);
    if (st != ORBX_OK) return st;
    const size_t pitch = align_up((size_t)width, 128), dpitch = align_up((size_t)width * 2, 128);
    if (!is_bgr) ORBX_CUDA(h, cudaMemcpy2DAsync(h->d_in, pitch, gray, step, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, h->stream));
    else {                                                          // cvtColor(BGR2GRAY) on the device, frontend.cpp:1084
        const size_t bpitch = align_up((size_t)width * 3, 128);
        if ((st = grow(h, &h->d_bgr, &h->bgr_cap, bpitch * height)) != ORBX_OK) return st;
        ORBX_CUDA(h, cudaMemcpy2DAsync(h->d_bgr, bpitch, gray, step, (size_t)width * 3, (size_t)height, cudaMemcpyHostToDevice, h->stream));
        launch_bgr2gray(h, h->d_bgr, bpitch, 0, h->d_in, pitch, 0, width, height, 1, h->stream);
    }
    const uint16_t *depth_zc = depth ? (const uint16_t *)mapped_device_view(depth) : nullptr;      // pinned depth is gathered in place
    if (depth && !depth_zc) ORBX_CUDA(h, cudaMemcpy2DAsync(h->d_depth_in, dpitch, depth, dstep, (size_t)width * 2, (size_t)height, cudaMemcpyHostToDevice, h->stream));
    if (nboxes > 0) {
        if (nboxes > h->boxes_cap) {
            cudaStreamSynchronize(h->stream);
            cudaFree(h->d_boxes); h->d_boxes = nullptr; h->boxes_cap = 0;
            ORBX_CUDA(h, cudaMalloc(&h->d_boxes, 2 * (size_t)nboxes * sizeof(orbx_box)));
            h->boxes_cap = nboxes;
        }
        ORBX_CUDA(h, cudaMemcpyAsync(h->d_boxes, boxes, (size_t)nboxes * sizeof(orbx_box), cudaMemcpyHostToDevice, h->stream));
    }
    const DevBoxes BX = { h->d_boxes, nullptr, 0, nboxes, drop_mask };
    st = run_pipeline(h, 1, h->d_in, pitch, 0, depth_zc ? depth_zc : (depth ? h->d_depth_in : nullptr), depth_zc ? dstep : dpitch, 0, BX,
                      h->d_kps_out, h->d_desc_out, h->max_kp, h->d_count_out);
    if (st != ORBX_OK) return st;
    // one packed D2H: [count | keypoints | descriptors] for min(cap, max_kp) entries
    const int ncopy = std::min(cap, h->max_kp);
    uint8_t *hb = h->h_out;
    ORBX_CUDA(h, cudaMemcpyAsync(hb, h->d_count_out, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (ncopy > 0) {
        ORBX_CUDA(h, cudaMemcpyAsync(hb + 64, h->d_kps_out, (size_t)ncopy * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(h, cudaMemcpyAsync(hb + 64 + (size_t)h->max_kp * sizeof(orbx_keypoint), h->d_desc_out, (size_t)ncopy * ORBX_DESC_BYTES, cudaMemcpyDeviceToHost, h->stream));
    }
    st = check_device_status(h);
    if (st != ORBX_OK) return st;
    const int n = *(int32_t *)hb;
    if (n > cap) { h->err = "output capacity too small"; *n_out = n; return ORBX_E_CAPACITY; }
    memcpy(kps, hb + 64, (size_t)n * sizeof(orbx_keypoint));
    memcpy(desc, hb + 64 + (size_t)h->max_kp * sizeof(orbx_keypoint), (size_t)n * ORBX_DESC_BYTES);
    *n_out = n;
    return ORBX_OK;
}

extern "C" orbx_status orbx_extract_filtered(orbx_handle *h, const uint8_t *gray, int32_t width, int32_t height, size_t step,
                                             const uint16_t *depth, size_t dstep, const orbx_box *boxes, int32_t nboxes, uint64_t drop_mask,
                                             orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *n_out)
{
    return extract_one(h, false, gray, width, height, step, depth, dstep, boxes, nboxes, drop_mask, kps, desc, cap, n_out);
}
extern "C" orbx_status orbx_extract_bgr(orbx_handle *h, const uint8_t *bgr, int32_t width, int32_t height, size_t step,
                                        const uint16_t *depth, size_t dstep, const orbx_box *boxes, int32_t nboxes, uint64_t drop_mask,
                                        orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *n_out)
{
    return extract_one(h, true, bgr, width, height, step, depth, dstep, boxes, nboxes, drop_mask, kps, desc, cap, n_out);
}
extern "C" orbx_status orbx_bgr2gray_device(orbx_handle *h, const uint8_t *d_bgr, int32_t nframes, int32_t width, int32_t height, size_t step,
                                            size_t fstride, uint8_t *d_gray, size_t gstep, size_t gfstride)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (!d_bgr || !d_gray || nframes < 1 || width < 1 || height < 1 || step < (size_t)width * 3 || gstep < (size_t)((width + 3) & ~3) || (gstep & 3) || ((uintptr_t)d_gray & 3)) {
        h->err = "bad bgr2gray arguments (gray rows must be 4-byte aligned and hold a multiple of 4 pixels)"; return ORBX_E_INVALID;
    }
    launch_bgr2gray(h, d_bgr, step, fstride, d_gray, gstep, gfstride, width, height, nframes, h->stream);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

extern "C" orbx_status orbx_extract(orbx_handle *h, const uint8_t *gray, int32_t width, int32_t height, size_t step,
                                    orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *n_out)
{
    return orbx_extract_filtered(h, gray, width, height, step, nullptr, 0, nullptr, 0, 0, kps, desc, cap, n_out);
}

// ---- stream step: extraction + depth filter + match against the previous frame (frontend.cpp:1094-1132) ----
extern "C" void orbx_track_reset(orbx_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaMemsetAsync(h->d_prev_count, 0, sizeof(int32_t), h->stream);
    h->prev_valid = 0;
}

static orbx_status track_matches(orbx_handle *h, int nframes, orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts,
                                 orbx_dmatch *d_matches, int32_t *d_mcounts, float max_dist);
static orbx_status track_chain(orbx_handle *h, const uint8_t *d_gray, int nframes, int width, int height, size_t step, size_t fstride,
                                const uint16_t *d_depth, size_t dstep, size_t dfstride, const DevBoxes &BX,
                                orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts,
                                orbx_dmatch *d_matches, int32_t *d_mcounts, float max_dist)
{
    if (cap > h->max_kp) { h->err = "cap_per_frame larger than the handle's max_keypoints"; return ORBX_E_INVALID; }
    orbx_status st = run_pipeline(h, nframes, d_gray, step, fstride, d_depth, dstep, dfstride, BX, d_kps, d_desc, cap, d_counts);
    if (st != ORBX_OK) return st;
    return track_matches(h, nframes, d_kps, d_desc, cap, d_counts, d_matches, d_mcounts, max_dist);
}

// matcher_.match(filtered, prev) + `distance < max_dist` for every frame of a batch, then prev = last frame (frontend.cpp:1123-1132, 1258-1259)
static orbx_status track_matches(orbx_handle *h, int nframes, orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts,
                                 orbx_dmatch *d_matches, int32_t *d_mcounts, float max_dist)
{
    (void)d_kps;
    // frame 0 against the carried state (prev count is 0 on the first frame => no matches, count 0)
    if (launch_match_core(h, d_desc, d_counts, cap, 0, h->d_prev_desc, h->d_prev_count, h->max_kp, 0, nullptr, nullptr, 1, 0,
                          1, max_dist, 0, d_matches, (size_t)cap, d_mcounts, nullptr) != 0) { h->err = "out of device memory (match scratch)"; return ORBX_E_NOMEM; }
    // frames 1..n-1 against their predecessor inside the batch: problem p = frame p+1
    if (nframes > 1 &&
        launch_match_core(h, d_desc + (size_t)cap * ORBX_DESC_BYTES, d_counts + 1, cap, (size_t)cap * ORBX_DESC_BYTES,
                          d_desc, d_counts, cap, (size_t)cap * ORBX_DESC_BYTES, nullptr, nullptr, nframes - 1, 0,
                          1, max_dist, 0, d_matches + cap, (size_t)cap, d_mcounts + 1, nullptr) != 0) { h->err = "out of device memory (match scratch)"; return ORBX_E_NOMEM; }
    // carry the last frame: prev_descriptors_ = filtered_descriptors.clone() (frontend.cpp:1258-1259)
    ORBX_CUDA(h, cudaMemcpyAsync(h->d_prev_desc, d_desc + (size_t)(nframes - 1) * cap * ORBX_DESC_BYTES, (size_t)cap * ORBX_DESC_BYTES, cudaMemcpyDeviceToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(h->d_prev_count, d_counts + (nframes - 1), sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    h->prev_valid = 1;
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}


// ---- ORBX_OPT_OVERLAP: two half-batches on two lanes (see OrbxLane) ----
static orbx_status ensure_alt_lane(orbx_handle *h)
{
    OrbxLane &A = h->alt;
    if (A.allocated) return ORBX_OK;
    const size_t B = (size_t)h->prm.max_batch;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    ORBX_CUDA(h, cudaStreamCreateWithPriority(&A.stream, cudaStreamNonBlocking, prio_hi));
    A.allocated = true;                                           // from here on orbx_destroy releases whatever exists
    ORBX_CUDA(h, cudaMalloc(&A.d_pyr, h->pyr_cap));
    ORBX_CUDA(h, cudaMalloc(&A.d_cand, h->cand_cap * sizeof(uint32_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_cand2, h->cand_cap * sizeof(uint32_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_qtmp, h->cand_cap * sizeof(uint32_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_owner, h->cand_cap * sizeof(uint16_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_owner2, h->cand_cap * sizeof(uint16_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_ncand, (B * ORBX_MAX_LEVELS + 4) * sizeof(int32_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_nsel, B * ORBX_MAX_LEVELS * sizeof(int32_t)));
    ORBX_CUDA(h, cudaMemset(A.d_nsel, 0, B * ORBX_MAX_LEVELS * sizeof(int32_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_sel, h->sel_cap * sizeof(uint32_t)));
    ORBX_CUDA(h, cudaMalloc(&A.d_kps_all, B * h->max_kp * sizeof(orbx_keypoint)));
    ORBX_CUDA(h, cudaMalloc(&A.d_desc_all, B * h->max_kp * ORBX_DESC_BYTES));
    ORBX_CUDA(h, cudaMalloc(&A.d_count_all, B * sizeof(int32_t)));
    A.d_mpart = nullptr; A.mpart_cap = 0; A.tmap_valid = false; A.tmap_l0 = nullptr;
    return ORBX_OK;
}

// exchange the handle's active scratch (and stream) with the other lane's
static void swap_lane(orbx_handle *h)
{
    OrbxLane &A = h->alt;
    std::swap(h->stream, A.stream);
    std::swap(h->d_pyr, A.d_pyr); std::swap(h->d_cand, A.d_cand); std::swap(h->d_cand2, A.d_cand2); std::swap(h->d_qtmp, A.d_qtmp);
    std::swap(h->d_owner, A.d_owner); std::swap(h->d_owner2, A.d_owner2); std::swap(h->d_ncand, A.d_ncand); std::swap(h->d_nsel, A.d_nsel);
    std::swap(h->d_sel, A.d_sel); std::swap(h->d_kps_all, A.d_kps_all); std::swap(h->d_desc_all, A.d_desc_all); std::swap(h->d_count_all, A.d_count_all);
    std::swap(h->d_mpart, A.d_mpart); std::swap(h->mpart_cap, A.mpart_cap);
    for (int l = 0; l < ORBX_MAX_LEVELS; l++) { std::swap(h->tmap[l], A.tmap[l]); std::swap(h->tmap_rz[l], A.tmap_rz[l]); std::swap(h->tmap_rz2[l], A.tmap_rz2[l]); std::swap(h->tmap_cell[l], A.tmap_cell[l]); }
    std::swap(h->tmap_valid, A.tmap_valid); std::swap(h->tmap_l0, A.tmap_l0); std::swap(h->tmap_l0_step, A.tmap_l0_step);
    std::swap(h->tmap_l0_fstride, A.tmap_l0_fstride); std::swap(h->tmap_l0_frames, A.tmap_l0_frames);
}

static bool overlap_applies(const orbx_handle *h, int nframes)
{
    return h->opt_overlap && nframes >= 32 && !h->opt_serial && !h->prof_on && h->opt_fused_blur;
}

// extraction (track == false) or the stream step (track == true) over nframes device frames: one chain, or two staggered half-batches
static orbx_status pipeline_device(orbx_handle *h, bool track, const uint8_t *d_gray, int nframes, int width, int height, size_t step, size_t fstride,
                                   const uint16_t *d_depth, size_t dstep, size_t dfstride, const DevBoxes &BX,
                                   orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts,
                                   orbx_dmatch *d_matches, int32_t *d_mcounts, float max_dist)
{
    if (!overlap_applies(h, nframes)) {
        if (track) return track_chain(h, d_gray, nframes, width, height, step, fstride, d_depth, dstep, dfstride, BX, d_kps, d_desc, cap, d_counts, d_matches, d_mcounts, max_dist);
        return run_pipeline(h, nframes, d_gray, step, fstride, d_depth, dstep, dfstride, BX, d_kps, d_desc, cap, d_counts);
    }
    orbx_status st = ensure_alt_lane(h);
    if (st != ORBX_OK) return st;
    const int nA = nframes / 2, nB = nframes - nA;
    h->in_overlap = true;
    // ---- first half on the handle's own lane; the second half may start once this half's pyramid is built ----
    h->ev_after_pyramid = h->ev_lane_fork;
    if (track) st = track_chain(h, d_gray, nA, width, height, step, fstride, d_depth, dstep, dfstride, BX, d_kps, d_desc, cap, d_counts, d_matches, d_mcounts, max_dist);
    else st = run_pipeline(h, nA, d_gray, step, fstride, d_depth, dstep, dfstride, BX, d_kps, d_desc, cap, d_counts);
    h->ev_after_pyramid = nullptr;
    if (st != ORBX_OK) { h->in_overlap = false; return st; }
    if (track) cudaEventRecord(h->ev_prev_ready, h->stream);                  // the first half's last frame is now the carried "previous frame"
    const int keep_batch = h->last_batch; const uint8_t *keep_l0 = h->last_l0; const size_t keep_step = h->last_l0_step, keep_fs = h->last_l0_fstride;
    // ---- second half on the other lane ----
    swap_lane(h);
    cudaStreamWaitEvent(h->stream, h->ev_lane_fork, 0);
    DevBoxes BB = BX;
    if (BB.off) BB.off += nA;
    const uint8_t *gB = d_gray + (size_t)nA * fstride;
    const uint16_t *dB = d_depth ? (const uint16_t *)((const uint8_t *)d_depth + (size_t)nA * dfstride) : nullptr;
    orbx_keypoint *kB = d_kps + (size_t)nA * cap; uint8_t *eB = d_desc + (size_t)nA * cap * ORBX_DESC_BYTES; int32_t *cB = d_counts + nA;
    if (track) {
        // the extraction of this half needs nothing from the first one; its frame 0 is matched against the first half's last frame
        st = run_pipeline(h, nB, gB, step, fstride, dB, dstep, dfstride, BB, kB, eB, cap, cB);
        if (st == ORBX_OK) { cudaStreamWaitEvent(h->stream, h->ev_prev_ready, 0); st = track_matches(h, nB, kB, eB, cap, cB, d_matches + (size_t)nA * cap, d_mcounts + nA, max_dist); }
    } else st = run_pipeline(h, nB, gB, step, fstride, dB, dstep, dfstride, BB, kB, eB, cap, cB);
    cudaEventRecord(h->ev_lane_join, h->stream);
    swap_lane(h);
    cudaStreamWaitEvent(h->stream, h->ev_lane_join, 0);                        // the handle's stream is complete when both halves are
    h->last_batch = keep_batch; h->last_l0 = keep_l0; h->last_l0_step = keep_step; h->last_l0_fstride = keep_fs;   // stage access: the first half
    h->in_overlap = false;
    if (st != ORBX_OK) return st;
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

extern "C" orbx_status orbx_track_batch_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                               int32_t width, int32_t height, size_t step, size_t frame_stride,
                                               const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                               orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap, int32_t *d_counts,
                                               orbx_dmatch *d_matches, int32_t *d_mcounts, float max_dist)
{
    return orbx_track_batch_boxes_device(h, d_gray, nframes, width, height, step, frame_stride, d_depth, dstep, dframe_stride,
                                         nullptr, nullptr, 0, 0, d_kps, d_desc, cap, d_counts, d_matches, d_mcounts, max_dist);
}

extern "C" orbx_status orbx_track_batch_boxes_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                                     int32_t width, int32_t height, size_t step, size_t frame_stride,
                                                     const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                                     const orbx_box *d_boxes, const int32_t *d_box_offsets, int32_t nboxes_total, uint64_t drop_mask,
                                                     orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap, int32_t *d_counts,
                                                     orbx_dmatch *d_matches, int32_t *d_mcounts, float max_dist)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (!d_gray || width <= 0 || height <= 0) { h->err = "empty image"; return ORBX_E_EMPTY; }
    if (nframes < 1 || nframes > h->prm.max_batch || !d_kps || !d_desc || !d_counts || !d_matches || !d_mcounts || cap < 1) { h->err = "bad batch arguments"; return ORBX_E_INVALID; }
    if (step < (size_t)width || (step & 15) || ((uintptr_t)d_gray & 15) || (frame_stride & 15)) { h->err = "device frames need 16-byte aligned base, step and frame stride"; return ORBX_E_INVALID; }
    if (d_depth && ((dstep & 1) || dstep < (size_t)width * 2)) { h->err = "bad depth step"; return ORBX_E_INVALID; }
    orbx_status st = set_geometry(h, width, height);
    if (st != ORBX_OK) return st;
    if (nboxes_total < 0 || (nboxes_total > 0 && (!d_boxes || !d_box_offsets))) { h->err = "per-frame boxes need the box array and nframes + 1 offsets"; return ORBX_E_INVALID; }
    const DevBoxes BX = { d_boxes, d_box_offsets, 0, nboxes_total, drop_mask };
    if (cap > h->max_kp) { h->err = "cap_per_frame larger than the handle's max_keypoints"; return ORBX_E_INVALID; }
    return pipeline_device(h, true, d_gray, nframes, width, height, step, frame_stride, d_depth, dstep, dframe_stride, BX, d_kps, d_desc, cap, d_counts, d_matches, d_mcounts, max_dist);
}

// Device view of a host buffer the GPU can read in place (pinned / registered, mapped under UVA); nullptr for pageable memory.
static const void *mapped_device_view(const void *host)
{
    cudaPointerAttributes a;
    if (!host || cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return a.devicePointer;
}

// Host-buffer batch driver.  Work is enqueued in CHUNKS of at most max_batch frames; a chunk owns one of two staging slots
// (input frames, output keypoints / descriptors / matches) and flows over three streams:
//     copy stream : H2D of the chunk's frames      (gray; depth only if it is pageable)
//     main stream : its kernels
//     out stream  : D2H of its results             (counts, keypoints, descriptors, matches, device status word)
// so the H2D of chunk i+1 and the D2H of chunk i-1 run under the kernels of chunk i.  The synchronous calls cut their frames
// into chunks of h->chunk frames; the *_submit / orbx_batch_wait pair enqueues one chunk per call and lets the CALLER keep two
// batches in flight (full-batch kernel efficiency, copies hidden).
// Depth maps in pinned host memory are NOT copied: the post-selection depth filter gathers its ~1000 samples per frame
// straight from the mapped host buffer (one 32-byte PCIe read each instead of 1.8 MB per frame).
struct BatchArgs {
    bool track; const uint8_t *gray; int width, height; size_t step; const uint16_t *depth; size_t dstep;
    orbx_keypoint *kps; uint8_t *desc; int cap; int32_t *counts; orbx_dmatch *matches; int32_t *mcounts; float max_dist;
    const orbx_box *boxes; const int32_t *box_off; uint64_t drop_mask;        // per-frame YOLO boxes (host), nullable
};

// `inl`: a call that is one small chunk (the single-frame latency path) keeps its copies on the kernels' stream — two cross-stream
// event hops less on the critical path; batches pipeline H2D / kernels / D2H over three streams
static orbx_status enqueue_chunk(orbx_handle *h, const BatchArgs &A, int f0, int nb, bool inl = false)
{
    const int slot = (int)(h->seq & 1);
    cudaStream_t cs = inl ? h->stream : h->copy_stream, os = inl ? h->stream : h->out_stream;
    const int width = A.width, height = A.height;
    const size_t pitch = align_up((size_t)width, 128), dpitch = align_up((size_t)width * 2, 128);
    const size_t fstride = pitch * height, dfstride = dpitch * height;
    const int SC = h->prm.max_batch;                              // slot capacity in frames
    const int kcap = std::min(A.cap, h->max_kp);
    const size_t slot_kp = (size_t)SC * h->max_kp;
    orbx_status st;
    orbx_dmatch *d_m = nullptr; int32_t *d_mc = nullptr;
    if (A.track) {
        if (h->mout_cap < 2 * slot_kp * sizeof(orbx_dmatch) + 2 * (size_t)SC * sizeof(int32_t)) {
            cudaStreamSynchronize(h->out_stream); cudaStreamSynchronize(h->stream);   // an older chunk may still be copying out of the old buffer
            if ((st = grow(h, (uint8_t **)&h->d_mout, &h->mout_cap, 2 * slot_kp * sizeof(orbx_dmatch) + 2 * (size_t)SC * sizeof(int32_t))) != ORBX_OK) return st;
        }
        d_m = h->d_mout + (size_t)slot * slot_kp;
        d_mc = (int32_t *)((uint8_t *)h->d_mout + 2 * slot_kp * sizeof(orbx_dmatch)) + slot * SC;
    }
    const uint8_t *depth_zc = A.depth ? (const uint8_t *)mapped_device_view(A.depth) : nullptr;
    const bool tight = (A.step == (size_t)width) && (pitch == (size_t)width);
    uint8_t *d_g = h->d_in + (size_t)slot * SC * fstride;
    uint8_t *d_d = (uint8_t *)h->d_depth_in + (size_t)slot * SC * dfstride;
    // ---- H2D (copy stream): the input slot is free once the kernels of the chunk that used it last are done ----
    ORBX_CUDA(h, cudaStreamWaitEvent(cs, h->ev_comp[slot], 0));
    if (tight) ORBX_CUDA(h, cudaMemcpyAsync(d_g, A.gray + (size_t)f0 * height * A.step, (size_t)nb * fstride, cudaMemcpyHostToDevice, cs));
    else for (int f = 0; f < nb; f++)
        ORBX_CUDA(h, cudaMemcpy2DAsync(d_g + (size_t)f * fstride, pitch, A.gray + (size_t)(f0 + f) * height * A.step, A.step,
                                       (size_t)width, (size_t)height, cudaMemcpyHostToDevice, cs));
    if (A.depth && !depth_zc) {
        if (A.dstep == dpitch) ORBX_CUDA(h, cudaMemcpyAsync(d_d, (const uint8_t *)A.depth + (size_t)f0 * height * A.dstep, (size_t)nb * dfstride, cudaMemcpyHostToDevice, cs));
        else for (int f = 0; f < nb; f++)
            ORBX_CUDA(h, cudaMemcpy2DAsync(d_d + (size_t)f * dfstride, dpitch, (const uint8_t *)A.depth + (size_t)(f0 + f) * height * A.dstep, A.dstep,
                                           (size_t)width * 2, (size_t)height, cudaMemcpyHostToDevice, cs));
    }
    // this chunk's YOLO boxes: the caller's offsets for frames f0..f0+nb as they are, its boxes [off[f0], off[f0+nb]) at the front of the slot
    DevBoxes BX = kNoBoxes;
    if (A.boxes && A.box_off) {
        const int b0 = A.box_off[f0], nbx = A.box_off[f0 + nb] - b0;
        if (nbx > 0) {
            if (nbx > h->boxes_cap) {                                  // rare: grow both slots (drains the pipeline)
                cudaStreamSynchronize(h->copy_stream); cudaStreamSynchronize(h->stream);
                cudaFree(h->d_boxes); h->d_boxes = nullptr; h->boxes_cap = 0;
                ORBX_CUDA(h, cudaMalloc(&h->d_boxes, 2 * (size_t)(nbx + 256) * sizeof(orbx_box)));
                h->boxes_cap = nbx + 256;
            }
            orbx_box *d_b = h->d_boxes + (size_t)slot * h->boxes_cap;
            int32_t *d_o = h->d_box_off + (size_t)slot * (SC + 1);
            ORBX_CUDA(h, cudaMemcpyAsync(d_b, A.boxes + b0, (size_t)nbx * sizeof(orbx_box), cudaMemcpyHostToDevice, cs));
            ORBX_CUDA(h, cudaMemcpyAsync(d_o, A.box_off + f0, (size_t)(nb + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
            BX.boxes = d_b; BX.off = d_o; BX.base = b0; BX.n = nbx; BX.drop_mask = A.drop_mask;
        }
    }
    ORBX_CUDA(h, cudaEventRecord(h->ev_in[slot], cs));
    // ---- kernels (main stream): the output slot is free once the D2H of the chunk that used it last is done ----
    ORBX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_in[slot], 0));
    ORBX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_out[slot], 0));
    orbx_keypoint *o_k = h->d_kps_out + (size_t)slot * slot_kp;
    uint8_t *o_d = h->d_desc_out + (size_t)slot * slot_kp * ORBX_DESC_BYTES;
    int32_t *o_c = h->d_count_out + slot * SC;
    const uint16_t *dd = nullptr; size_t dds = 0, ddf = 0;
    if (depth_zc) { dd = (const uint16_t *)(depth_zc + (size_t)f0 * height * A.dstep); dds = A.dstep; ddf = (size_t)height * A.dstep; }
    else if (A.depth) { dd = (const uint16_t *)d_d; dds = dpitch; ddf = dfstride; }
    // this chunk's kernels report device-side capacity flags into the slot's own word, so a flag raised by one chunk is never
    // attributed to the other chunk in flight or to a later call
    h->d_status = h->d_status_base + 1 + slot;           // (zeroed at the start of the call: host_batch / submit_batch)
    st = pipeline_device(h, A.track, d_g, nb, width, height, pitch, fstride, dd, dds, ddf, BX, o_k, o_d, h->max_kp, o_c, d_m, d_mc, A.max_dist);
    int32_t *slot_status = h->d_status;
    h->d_status = h->d_status_base;
    if (st != ORBX_OK) return st;
    ORBX_CUDA(h, cudaEventRecord(h->ev_comp[slot], h->stream));
    // ---- D2H (out stream) ----
    ORBX_CUDA(h, cudaStreamWaitEvent(os, h->ev_comp[slot], 0));
    ORBX_CUDA(h, cudaMemcpyAsync(A.counts + f0, o_c, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, os));
    ORBX_CUDA(h, cudaMemcpy2DAsync(A.kps + (size_t)f0 * A.cap, (size_t)A.cap * sizeof(orbx_keypoint), o_k, (size_t)h->max_kp * sizeof(orbx_keypoint),
                                   (size_t)kcap * sizeof(orbx_keypoint), (size_t)nb, cudaMemcpyDeviceToHost, os));
    ORBX_CUDA(h, cudaMemcpy2DAsync(A.desc + (size_t)f0 * A.cap * ORBX_DESC_BYTES, (size_t)A.cap * ORBX_DESC_BYTES, o_d, (size_t)h->max_kp * ORBX_DESC_BYTES,
                                   (size_t)kcap * ORBX_DESC_BYTES, (size_t)nb, cudaMemcpyDeviceToHost, os));
    if (A.track) {
        ORBX_CUDA(h, cudaMemcpyAsync(A.mcounts + f0, d_mc, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, os));
        ORBX_CUDA(h, cudaMemcpy2DAsync(A.matches + (size_t)f0 * A.cap, (size_t)A.cap * sizeof(orbx_dmatch), d_m, (size_t)h->max_kp * sizeof(orbx_dmatch),
                                       (size_t)kcap * sizeof(orbx_dmatch), (size_t)nb, cudaMemcpyDeviceToHost, os));
    }
    ORBX_CUDA(h, cudaMemcpyAsync(h->h_status + 1 + slot, slot_status, sizeof(int32_t), cudaMemcpyDeviceToHost, os));
    ORBX_CUDA(h, cudaEventRecord(h->ev_out[slot], os));
    h->seq++;
    return ORBX_OK;
}

// completion of the chunk in `slot`: results are in the caller's buffers; device capacity flags and output capacity are reported here
static orbx_status finish_slot(orbx_handle *h, int slot, const int32_t *counts, int nframes, int cap)
{
    ORBX_CUDA(h, cudaEventSynchronize(h->ev_out[slot]));
    const int s = h->h_status[1 + slot];
    if (s != 0) {
        h->h_status[1 + slot] = 0;
        if (s & ORBX_DS_INTERNAL) { h->err = "internal: a frame of the dense FAST kernel did not complete within its wait bound; results of this call are incomplete"; return ORBX_E_CUDA; }
        h->err = std::string("device capacity exceeded:") + ((s & ORBX_DS_CAND_OVERFLOW) ? " candidate list (lower cand_divisor)" : "") +
                 ((s & ORBX_DS_NODE_OVERFLOW) ? " quadtree nodes" : "") + ((s & ORBX_DS_KP_OVERFLOW) ? " keypoint output (raise cap / max_keypoints)" : "");
        return ORBX_E_CAPACITY;
    }
    for (int f = 0; f < nframes; f++) if (counts[f] > cap) { h->err = "output capacity too small"; return ORBX_E_CAPACITY; }
    return ORBX_OK;
}

static orbx_status drain_all(orbx_handle *h)
{
    cudaStreamSynchronize(h->copy_stream); cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->out_stream);
    h->pending[0].active = h->pending[1].active = false;
    return ORBX_OK;
}

static orbx_status host_batch(orbx_handle *h, const BatchArgs &A, int nframes)
{
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    orbx_status st = set_geometry(h, A.width, A.height);
    if (st != ORBX_OK) return st;
    const int C = h->chunk;
    const bool inl = nframes <= C && nframes <= 4;
    ORBX_CUDA(h, cudaMemsetAsync(h->d_status_base + 1, 0, 2 * sizeof(int32_t), h->stream));   // flags stay sticky across the chunks of THIS call only
    for (int f0 = 0; f0 < nframes; f0 += C) {
        const int nb = std::min(C, nframes - f0);
        if ((st = enqueue_chunk(h, A, f0, nb, inl)) != ORBX_OK) { drain_all(h); return st; }
    }
    ORBX_CUDA(h, cudaStreamSynchronize(inl ? h->stream : h->out_stream));
    // both slots are always inspected (and their flags consumed), so nothing raised by this call survives into the next one
    const orbx_status s0 = finish_slot(h, 0, A.counts, nframes, A.cap), s1 = finish_slot(h, 1, A.counts, 0, A.cap);
    return s0 != ORBX_OK ? s0 : s1;
}

static orbx_status submit_batch(orbx_handle *h, const BatchArgs &A, int nframes, int32_t *ticket)
{
    if (!ticket) { h->err = "null ticket"; return ORBX_E_INVALID; }
    if (nframes < 1 || nframes > h->prm.max_batch) { h->err = "an asynchronous submission holds 1..max_batch frames"; return ORBX_E_INVALID; }
    const int slot = (int)(h->seq & 1);
    if (h->pending[slot].active) { h->err = "two batches are already in flight: wait for the older ticket first"; return ORBX_E_INVALID; }
    orbx_status st = set_geometry(h, A.width, A.height);      // (a geometry change synchronises the streams itself)
    if (st != ORBX_OK) return st;
    const uint64_t t = h->seq;
    ORBX_CUDA(h, cudaMemsetAsync(h->d_status_base + 1 + slot, 0, sizeof(int32_t), h->stream));
    if ((st = enqueue_chunk(h, A, 0, nframes)) != ORBX_OK) { drain_all(h); return st; }
    h->pending[slot].active = true; h->pending[slot].ticket = t; h->pending[slot].counts = A.counts; h->pending[slot].nframes = nframes; h->pending[slot].cap = A.cap;
    *ticket = (int32_t)(t & 0x7FFFFFFF);
    return ORBX_OK;
}

extern "C" orbx_status orbx_batch_wait(orbx_handle *h, int32_t ticket)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    for (int slot = 0; slot < 2; slot++) {
        if (h->pending[slot].active && (int32_t)(h->pending[slot].ticket & 0x7FFFFFFF) == ticket) {
            // tickets complete in submission order: the other slot's older ticket, if any, must be collected first
            const int o = slot ^ 1;
            if (h->pending[o].active && h->pending[o].ticket < h->pending[slot].ticket) { h->err = "wait for the older ticket first"; return ORBX_E_INVALID; }
            h->pending[slot].active = false;
            return finish_slot(h, slot, h->pending[slot].counts, h->pending[slot].nframes, h->pending[slot].cap);
        }
    }
    h->err = "unknown ticket";
    return ORBX_E_INVALID;
}

static bool batch_args_ok(orbx_handle *h, const BatchArgs &A, int nframes)
{
    if (!A.gray || A.width <= 0 || A.height <= 0) { h->err = "empty image"; return false; }
    if (nframes < 0 || !A.kps || !A.desc || !A.counts || A.cap < 1 || A.step < (size_t)A.width ||
        (A.track && (!A.matches || !A.mcounts))) { h->err = "bad arguments"; return false; }
    if (A.depth && (A.dstep < (size_t)A.width * 2 || (A.dstep & 1) || ((uintptr_t)A.depth & 1))) { h->err = "bad depth buffer: rows of at least 2*width bytes, even step, 2-byte aligned"; return false; }
    if ((A.boxes != nullptr) != (A.box_off != nullptr)) { h->err = "per-frame boxes need both the box array and nframes + 1 offsets"; return false; }
    if (A.box_off) {
        if (A.box_off[0] < 0) { h->err = "box offsets must start at >= 0"; return false; }
        for (int f = 0; f < nframes; f++) if (A.box_off[f + 1] < A.box_off[f]) { h->err = "box offsets must not decrease"; return false; }
    }
    return true;
}

// the five host-buffer batch entry points share one body: validate, then run synchronously (chunk pipeline) or enqueue one batch
static orbx_status host_entry(orbx_handle *h, const BatchArgs &A, int nframes, int32_t *ticket, bool async)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (!A.gray || A.width <= 0 || A.height <= 0) { h->err = "empty image"; return ORBX_E_EMPTY; }
    if (!batch_args_ok(h, A, nframes)) return ORBX_E_INVALID;
    return async ? submit_batch(h, A, nframes, ticket) : host_batch(h, A, nframes);
}

extern "C" orbx_status orbx_extract_batch(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                          size_t step, const uint16_t *depth, size_t dstep,
                                          orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *counts)
{
    const BatchArgs A = { false, gray, width, height, step, depth, dstep, kps, desc, cap, counts, nullptr, nullptr, 0.f, nullptr, nullptr, 0 };
    return host_entry(h, A, nframes, nullptr, false);
}
extern "C" orbx_status orbx_extract_batch_boxes(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                                size_t step, const uint16_t *depth, size_t dstep,
                                                const orbx_box *boxes, const int32_t *box_offsets, uint64_t drop_mask,
                                                orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *counts)
{
    const BatchArgs A = { false, gray, width, height, step, depth, dstep, kps, desc, cap, counts, nullptr, nullptr, 0.f, boxes, box_offsets, drop_mask };
    return host_entry(h, A, nframes, nullptr, false);
}
extern "C" orbx_status orbx_track_batch(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                        size_t step, const uint16_t *depth, size_t dstep,
                                        orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *counts,
                                        orbx_dmatch *matches, int32_t *mcounts, float max_dist)
{
    const BatchArgs A = { true, gray, width, height, step, depth, dstep, kps, desc, cap, counts, matches, mcounts, max_dist, nullptr, nullptr, 0 };
    return host_entry(h, A, nframes, nullptr, false);
}
extern "C" orbx_status orbx_track_batch_boxes(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                              size_t step, const uint16_t *depth, size_t dstep,
                                              const orbx_box *boxes, const int32_t *box_offsets, uint64_t drop_mask,
                                              orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *counts,
                                              orbx_dmatch *matches, int32_t *mcounts, float max_dist)
{
    const BatchArgs A = { true, gray, width, height, step, depth, dstep, kps, desc, cap, counts, matches, mcounts, max_dist, boxes, box_offsets, drop_mask };
    return host_entry(h, A, nframes, nullptr, false);
}
extern "C" orbx_status orbx_extract_batch_submit(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                                 size_t step, const uint16_t *depth, size_t dstep,
                                                 orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *counts, int32_t *ticket)
{
    const BatchArgs A = { false, gray, width, height, step, depth, dstep, kps, desc, cap, counts, nullptr, nullptr, 0.f, nullptr, nullptr, 0 };
    return host_entry(h, A, nframes, ticket, true);
}
extern "C" orbx_status orbx_track_batch_submit(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                               size_t step, const uint16_t *depth, size_t dstep,
                                               orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *counts,
                                               orbx_dmatch *matches, int32_t *mcounts, float max_dist, int32_t *ticket)
{
    const BatchArgs A = { true, gray, width, height, step, depth, dstep, kps, desc, cap, counts, matches, mcounts, max_dist, nullptr, nullptr, 0 };
    return host_entry(h, A, nframes, ticket, true);
}
extern "C" orbx_status orbx_track_batch_boxes_submit(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                                     size_t step, const uint16_t *depth, size_t dstep,
                                                     const orbx_box *boxes, const int32_t *box_offsets, uint64_t drop_mask,
                                                     orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *counts,
                                                     orbx_dmatch *matches, int32_t *mcounts, float max_dist, int32_t *ticket)
{
    const BatchArgs A = { true, gray, width, height, step, depth, dstep, kps, desc, cap, counts, matches, mcounts, max_dist, boxes, box_offsets, drop_mask };
    return host_entry(h, A, nframes, ticket, true);
}

// ---- stage access (mvImagePyramid is public in the reference, ORBextractor.hpp:84) ----
extern "C" orbx_status orbx_get_pyramid_level(orbx_handle *h, int32_t frame, int32_t level, uint8_t *out, size_t out_step)
{
    if (!h || !out || level < 0 || level >= h->geo.nlevels || frame < 0 || frame >= h->last_batch) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (h->cv) {                                                   // profile C keeps the levels of the LAST frame it processed
        if (frame != h->last_batch - 1) { h->err = "profile C: stage access is for the last frame of the batch"; return ORBX_E_INVALID; }
        return orbx_cvorb_get_level(h, level, 0, out, out_step, h->last_l0 + (size_t)frame * h->last_l0_fstride, h->last_l0_step);
    }
    const LevelGeom &g = h->geo.lv[level];
    const uint8_t *src; size_t step;
    if (level == 0) { src = h->last_l0 + (size_t)frame * h->last_l0_fstride; step = h->last_l0_step; }
    else { src = h->d_pyr + (size_t)frame * h->pyr_slab + g.off; step = g.pitch; }
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBX_CUDA(h, cudaMemcpy2D(out, out_step, src, step, (size_t)g.w, (size_t)g.h, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}
extern "C" orbx_status orbx_get_blurred_level(orbx_handle *h, int32_t frame, int32_t level, uint8_t *out, size_t out_step)
{
    if (!h || !out || level < 0 || level >= h->geo.nlevels || frame < 0 || frame >= h->last_batch) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (h->cv) {
        if (frame != h->last_batch - 1) { h->err = "profile C: stage access is for the last frame of the batch"; return ORBX_E_INVALID; }
        return orbx_cvorb_get_level(h, level, 1, out, out_step, nullptr, 0);
    }
    const LevelGeom &g = h->geo.lv[level];
    if (!h->blur_valid) {                                                          // fused mode: materialise the last batch's blurred levels now
        if (launch_blur(h, h->last_batch, h->last_l0, h->last_l0_step, h->last_l0_fstride, h->stream) != 0) { h->err = "cuTensorMapEncodeTiled failed"; return ORBX_E_CUDA; }
        h->blur_valid = true;
    }
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBX_CUDA(h, cudaMemcpy2D(out, out_step, h->d_blur + (size_t)frame * h->blur_slab + g.boff, g.bpitch, (size_t)g.w, (size_t)g.h, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}
// dense FAST formulation: the iniThFAST score map of a level (S - 1 where S > iniThFAST, else 0 — cv::FAST's score buffer over the whole
// level), valid after a batch that ran k_fast_dense.  out: (h - 38) rows x (w - 38) columns, level pixel (19 + c, 19 + r) at out[r][c];
// pixels outside the cell grid read 0.
extern "C" orbx_status orbx_get_fast_scores(orbx_handle *h, int32_t frame, int32_t level, uint8_t *out, size_t out_step)
{
    if (!h || !out || level < 0 || level >= h->geo.nlevels || frame < 0 || frame >= h->last_batch) return ORBX_E_INVALID;
    if (h->cv || !h->dense_ok || h->dgeo.ntiles <= 0) { h->err = "no dense FAST run on this handle"; return ORBX_E_UNSUPPORTED; }
    cudaSetDevice(h->device);
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    const LevelGeom &g = h->geo.lv[level];
    const DenseLevel &d = h->dgeo.lv[level];
    const int ow = g.w - 2 * (ORBX_BORDER + 3), oh = g.h - 2 * (ORBX_BORDER + 3);
    if (ow <= 0 || oh <= 0 || out_step < (size_t)ow) return ORBX_E_INVALID;
    for (int r = 0; r < oh; r++) memset(out + (size_t)r * out_step, 0, (size_t)ow);
    if (d.ntx <= 0) return ORBX_OK;
    const int rows = std::min(oh, d.nty * 16), cols = std::min(ow, d.map_pitch - (ORBX_BORDER + 3 - 4));
    ORBX_CUDA(h, cudaMemcpy2D(out, out_step, h->d_smap + (size_t)frame * h->dgeo.map_bytes + d.map_off + (ORBX_BORDER + 3 - 4), (size_t)d.map_pitch,
                              (size_t)cols, (size_t)rows, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

// dense FAST formulation, stage access for tests: the corners k_fast_dense left to k_fast_nms (those on tile edges, and every corner of a
// tile with more pre-test survivors than its queue holds) of frame slot `frame`, as (x, y, score) triples like orbx_get_candidates
extern "C" orbx_status orbx_get_fast_edge_corners(orbx_handle *h, int32_t frame, int32_t level, int32_t *out_xys, int32_t cap, int32_t *n_out)
{
    if (!h || !out_xys || !n_out || level < 0 || level >= h->geo.nlevels || frame < 0 || frame >= h->last_batch) return ORBX_E_INVALID;
    if (h->cv || !h->dense_ok || h->dgeo.ntiles <= 0) { h->err = "no dense FAST run on this handle"; return ORBX_E_UNSUPPORTED; }
    cudaSetDevice(h->device);
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    const DenseLevel &d = h->dgeo.lv[level];
    int32_t n = 0;
    ORBX_CUDA(h, cudaMemcpy(&n, h->d_dense_zero + 4 + frame * h->geo.nlevels + level, sizeof(int32_t), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > d.cl_cap) { h->err = "corner list overflowed"; return ORBX_E_CAPACITY; }
    if (n > cap) return ORBX_E_CAPACITY;
    std::vector<uint32_t> tmp((size_t)std::max(n, 1));
    ORBX_CUDA(h, cudaMemcpy(tmp.data(), h->d_clist + (size_t)frame * h->dgeo.cl_entries + d.cl_off, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) { out_xys[3 * i] = orbx_px(tmp[i]); out_xys[3 * i + 1] = orbx_py(tmp[i]); out_xys[3 * i + 2] = orbx_ps(tmp[i]); }
    return ORBX_OK;
}

extern "C" orbx_status orbx_get_candidates(orbx_handle *h, int32_t frame, int32_t level, int32_t *out_xys, int32_t cap, int32_t *n_out)
{
    if (!h || !out_xys || !n_out || level < 0 || level >= h->geo.nlevels || frame < 0 || frame >= h->last_batch) return ORBX_E_INVALID;
    if (h->cv) { h->err = "candidate lists are a profile-S stage"; return ORBX_E_UNSUPPORTED; }
    cudaSetDevice(h->device);
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    const LevelGeom &g = h->geo.lv[level];
    int32_t n = 0;
    ORBX_CUDA(h, cudaMemcpy(&n, h->d_ncand + frame * h->geo.nlevels + level, sizeof(int32_t), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > g.cand_cap) { h->err = "candidate list overflowed"; return ORBX_E_CAPACITY; }
    if (n > cap) return ORBX_E_CAPACITY;
    // the quadtree ping-pongs between d_cand and d_cand2 but only permutes inside the list: either buffer holds the full set
    std::vector<uint32_t> tmp((size_t)std::max(n, 1));
    ORBX_CUDA(h, cudaMemcpy(tmp.data(), h->d_cand + (size_t)frame * h->geo.cand_entries + g.cand_off, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) { out_xys[3 * i] = orbx_px(tmp[i]); out_xys[3 * i + 1] = orbx_py(tmp[i]); out_xys[3 * i + 2] = orbx_ps(tmp[i]); }
    return ORBX_OK;
}
extern "C" orbx_status orbx_harris_responses(orbx_handle *h, int32_t frame, int32_t level, const int32_t *xy, int32_t n, int32_t bs, float k, float *out)
{
    if (!h || level < 0 || level >= h->geo.nlevels || frame < 0 || frame >= h->last_batch || n < 0 || (n > 0 && (!xy || !out)) || bs < 1 || bs > 31 || !(bs & 1)) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n == 0) return ORBX_OK;
    const LevelGeom &g = h->geo.lv[level];
    const uint8_t *src; size_t step;
    if (level == 0) { src = h->last_l0 + (size_t)frame * h->last_l0_fstride; step = h->last_l0_step; }
    else { src = h->d_pyr + (size_t)frame * h->pyr_slab + g.off; step = g.pitch; }
    orbx_status st;
    if ((st = grow(h, &h->d_mq, &h->mq_cap, (size_t)n * 12)) != ORBX_OK) return st;
    int32_t *d_xy = (int32_t *)h->d_mq; float *d_o = (float *)(h->d_mq + (size_t)n * 8);
    ORBX_CUDA(h, cudaMemcpyAsync(d_xy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    launch_harris(h, src, step, g.w, g.h, d_xy, n, bs, k, d_o);
    ORBX_CUDA(h, cudaMemcpyAsync(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}
extern "C" orbx_status orbx_get_level_counts(orbx_handle *h, int32_t frame, int32_t *out)
{
    if (!h || !out || frame < 0 || frame >= h->last_batch) return ORBX_E_INVALID;
    if (h->cv) { h->err = "per-level counts: read the octave field of the keypoints (profile C)"; return ORBX_E_UNSUPPORTED; }
    cudaSetDevice(h->device);
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBX_CUDA(h, cudaMemcpy(out, h->d_nsel + frame * h->geo.nlevels, sizeof(int32_t) * h->geo.nlevels, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

// ---- matching ----
static int ratio_to_num(float ratio) { return ratio > 0.f ? (int)lrintf(ratio * 1024.f) : 0; }

extern "C" orbx_status orbx_match_device(orbx_handle *h, const uint8_t *d_q, int32_t nq, const uint8_t *d_t, int32_t nt,
                                         int32_t k, float max_dist, float ratio, orbx_dmatch *d_out, int32_t *d_n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (nq < 0 || nt < 0 || (k != 1 && k != 2) || !d_out || (nq > 0 && !d_q) || (nt > 0 && !d_t)) { h->err = "bad match arguments"; return ORBX_E_INVALID; }
    if (nq == 0) { if (d_n_out) ORBX_CUDA(h, cudaMemsetAsync(d_n_out, 0, sizeof(int32_t), h->stream)); return ORBX_OK; }
    if (launch_match_core(h, d_q, nullptr, nq, 0, d_t, nullptr, nt, 0, nullptr, nullptr, 1, 0, k, max_dist, ratio_to_num(ratio),
                          d_out, (size_t)nq * k, d_n_out, nullptr) != 0) { h->err = "out of device memory (match scratch)"; return ORBX_E_NOMEM; }
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

static orbx_status grow(orbx_handle *h, uint8_t **p, size_t *cap, size_t need)
{
    if (need <= *cap) return ORBX_OK;
    if (*p) { cudaStreamSynchronize(h->stream); cudaFree(*p); *p = nullptr; *cap = 0; }
    ORBX_CUDA(h, cudaMalloc(p, need));
    *cap = need;
    return ORBX_OK;
}

extern "C" orbx_status orbx_match(orbx_handle *h, const uint8_t *q, int32_t nq, const uint8_t *t, int32_t nt,
                                  int32_t k, float max_dist, float ratio, orbx_dmatch *out, int32_t *n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n_out) *n_out = 0;
    if (nq < 0 || nt < 0 || (k != 1 && k != 2) || !out || !n_out || (nq > 0 && !q) || (nt > 0 && !t)) { h->err = "bad match arguments"; return ORBX_E_INVALID; }
    if (nq == 0) return ORBX_OK;
    orbx_status st;
    if ((st = grow(h, &h->d_mq, &h->mq_cap, (size_t)nq * ORBX_DESC_BYTES)) != ORBX_OK) return st;
    if ((st = grow(h, &h->d_mt, &h->mt_cap, (size_t)std::max(nt, 1) * ORBX_DESC_BYTES)) != ORBX_OK) return st;
    if ((st = grow(h, (uint8_t **)&h->d_mout, &h->mout_cap, (size_t)nq * k * sizeof(orbx_dmatch))) != ORBX_OK) return st;
    ORBX_CUDA(h, cudaMemcpyAsync(h->d_mq, q, (size_t)nq * ORBX_DESC_BYTES, cudaMemcpyHostToDevice, h->stream));
    if (nt > 0) ORBX_CUDA(h, cudaMemcpyAsync(h->d_mt, t, (size_t)nt * ORBX_DESC_BYTES, cudaMemcpyHostToDevice, h->stream));
    st = orbx_match_device(h, h->d_mq, nq, h->d_mt, nt, k, max_dist, ratio, h->d_mout, h->d_mcount);
    if (st != ORBX_OK) return st;
    int32_t n = 0;
    ORBX_CUDA(h, cudaMemcpyAsync(&n, h->d_mcount, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    { const orbx_status ds = check_device_status(h); if (ds != ORBX_OK) return ds; }      // synchronises; a broken tensor-memory pipeline (never seen) must not return silently
    if (n > 0) ORBX_CUDA(h, cudaMemcpy(out, h->d_mout, (size_t)n * sizeof(orbx_dmatch), cudaMemcpyDeviceToHost));
    *n_out = n;
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_pairs_device(orbx_handle *h, const uint8_t *d_desc, const int32_t *d_counts, int32_t cap,
                                               const int32_t *q_frame, const int32_t *t_frame, int32_t npairs,
                                               int32_t k, float max_dist, float ratio, orbx_dmatch *d_out, int32_t *d_n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (!d_desc || !d_counts || cap < 1 || npairs < 0 || (k != 1 && k != 2) || !d_out || !d_n_out || (npairs > 0 && (!q_frame || !t_frame))) { h->err = "bad match arguments"; return ORBX_E_INVALID; }
    if (npairs == 0) return ORBX_OK;
    orbx_status st;
    size_t need = (size_t)npairs * 2 * sizeof(int32_t);
    if ((st = grow(h, &h->d_mq, &h->mq_cap, need)) != ORBX_OK) return st;     // reuse the query scratch for the pair lists
    int32_t *d_qsel = (int32_t *)h->d_mq, *d_tsel = d_qsel + npairs;
    ORBX_CUDA(h, cudaMemcpyAsync(d_qsel, q_frame, (size_t)npairs * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(d_tsel, t_frame, (size_t)npairs * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (launch_match_core(h, d_desc, d_counts, cap, (size_t)cap * ORBX_DESC_BYTES, d_desc, d_counts, cap, (size_t)cap * ORBX_DESC_BYTES,
                          d_qsel, d_tsel, npairs, 0, k, max_dist, ratio_to_num(ratio), d_out, (size_t)cap * k, d_n_out, nullptr) != 0) {
        h->err = "out of device memory (match scratch)"; return ORBX_E_NOMEM;
    }
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

// ---- landmark database (Backend::associateObservation, descriptor stage) ----
extern "C" orbx_status orbx_db_create(orbx_handle *h, int64_t capacity_rows, uint32_t first_index, orbx_db **out)
{
    if (!h || !out || capacity_rows < 1) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    orbx_db *db = new orbx_db();
    memset(db, 0, sizeof(*db));
    db->h = h; db->cap = capacity_rows; db->rows = 0; db->first_index = first_index;
    if (cudaMalloc(&db->d_rows, (size_t)capacity_rows * ORBX_DESC_BYTES) != cudaSuccess) { delete db; h->err = "out of device memory (db)"; return ORBX_E_NOMEM; }
    *out = db;
    return ORBX_OK;
}
extern "C" void orbx_db_destroy(orbx_db *db)
{
    if (!db) return;
    cudaSetDevice(db->h->device);
    cudaStreamSynchronize(db->h->stream);
    if (db->d_rows) cudaFree(db->d_rows);
    if (db->d_q) cudaFree(db->d_q);
    if (db->d_out) cudaFree(db->d_out);
    if (db->d_pos) cudaFree(db->d_pos);
    if (db->d_qpx) cudaFree(db->d_qpx);
    delete db;
}
extern "C" int64_t orbx_db_rows(const orbx_db *db) { return db ? db->rows : 0; }
extern "C" orbx_status orbx_db_append(orbx_db *db, const uint8_t *rows, int64_t n)
{
    if (!db || n < 0 || (n > 0 && !rows)) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    if (db->rows + n > db->cap) { h->err = "database capacity exceeded"; return ORBX_E_CAPACITY; }
    ORBX_CUDA(h, cudaMemcpyAsync(db->d_rows + (size_t)db->rows * ORBX_DESC_BYTES, rows, (size_t)n * ORBX_DESC_BYTES, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    db->rows += n;
    return ORBX_OK;
}
extern "C" orbx_status orbx_db_append_device(orbx_db *db, const uint8_t *d_rows, int64_t n)
{
    if (!db || n < 0 || (n > 0 && !d_rows)) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    if (db->rows + n > db->cap) { h->err = "database capacity exceeded"; return ORBX_E_CAPACITY; }
    ORBX_CUDA(h, cudaMemcpyAsync(db->d_rows + (size_t)db->rows * ORBX_DESC_BYTES, d_rows, (size_t)n * ORBX_DESC_BYTES, cudaMemcpyDeviceToDevice, h->stream));
    db->rows += n;
    return ORBX_OK;
}
extern "C" orbx_status orbx_db_query_top2_device(orbx_db *db, const uint8_t *d_q, int32_t nq, orbx_top2 *d_out)
{
    if (!db || nq < 0 || !d_out || (nq > 0 && !d_q)) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    if (nq == 0) return ORBX_OK;
    if (db->rows > 0x7FFFFFFF) { h->err = "shard too large"; return ORBX_E_UNSUPPORTED; }
    if (launch_match_core(h, d_q, nullptr, nq, 0, db->d_rows, nullptr, (int)db->rows, 0, nullptr, nullptr, 1, db->first_index,
                          2, 0.f, 0, nullptr, 0, nullptr, d_out) != 0) { h->err = "out of device memory (match scratch)"; return ORBX_E_NOMEM; }
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
static orbx_status db_stage_queries(orbx_db *db, const uint8_t *q, int32_t nq, size_t out_bytes)
{
    orbx_handle *h = db->h;
    if ((size_t)nq * ORBX_DESC_BYTES > db->q_cap) {
        if (db->d_q) { cudaStreamSynchronize(h->stream); cudaFree(db->d_q); db->d_q = nullptr; db->q_cap = 0; }
        ORBX_CUDA(h, cudaMalloc(&db->d_q, (size_t)nq * ORBX_DESC_BYTES)); db->q_cap = (size_t)nq * ORBX_DESC_BYTES;
    }
    if (out_bytes > db->out_cap) {
        if (db->d_out) { cudaStreamSynchronize(h->stream); cudaFree(db->d_out); db->d_out = nullptr; db->out_cap = 0; }
        ORBX_CUDA(h, cudaMalloc(&db->d_out, out_bytes)); db->out_cap = out_bytes;
    }
    ORBX_CUDA(h, cudaMemcpyAsync(db->d_q, q, (size_t)nq * ORBX_DESC_BYTES, cudaMemcpyHostToDevice, h->stream));
    return ORBX_OK;
}
extern "C" orbx_status orbx_db_query_top2(orbx_db *db, const uint8_t *q, int32_t nq, orbx_top2 *out)
{
    if (!db || nq < 0 || !out || (nq > 0 && !q)) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    if (nq == 0) return ORBX_OK;
    orbx_status st = db_stage_queries(db, q, nq, (size_t)nq * sizeof(orbx_top2));
    if (st != ORBX_OK) return st;
    if ((st = orbx_db_query_top2_device(db, db->d_q, nq, db->d_out)) != ORBX_OK) return st;
    ORBX_CUDA(h, cudaMemcpyAsync(out, db->d_out, (size_t)nq * sizeof(orbx_top2), cudaMemcpyDeviceToHost, h->stream));
    { const orbx_status ds = check_device_status(h); if (ds != ORBX_OK) return ds; }      // synchronises; a broken tensor-memory pipeline (never seen) must not return silently
    return ORBX_OK;
}
extern "C" orbx_status orbx_merge_top2_device(orbx_handle *h, const orbx_top2 *d_parts, int32_t nshards, int32_t nq, orbx_top2 *d_out)
{
    if (!h || !d_parts || !d_out || nshards < 1 || nq < 0) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    launch_merge_top2(h, d_parts, nshards, nq, d_out);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
extern "C" orbx_status orbx_db_query_radius(orbx_db *db, const uint8_t *q, int32_t nq, float max_dist,
                                            orbx_dmatch *out, int32_t cap, int32_t *n_out)
{
    if (!db || nq < 0 || !out || !n_out || cap < 0 || (nq > 0 && !q)) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    *n_out = 0;
    if (nq == 0 || db->rows == 0) return ORBX_OK;
    orbx_status st = db_stage_queries(db, q, nq, (size_t)std::max(cap, 1) * sizeof(orbx_dmatch));
    if (st != ORBX_OK) return st;
    ORBX_CUDA(h, cudaMemsetAsync(h->d_mcount, 0, sizeof(int32_t), h->stream));
    launch_match_radius(h, db->d_q, nq, db->d_rows, (int)db->rows, db->first_index, max_dist, (orbx_dmatch *)db->d_out, cap, h->d_mcount);
    int32_t n = 0;
    ORBX_CUDA(h, cudaMemcpyAsync(&n, h->d_mcount, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    *n_out = n;
    if (n > cap) { h->err = "radius output capacity too small"; return ORBX_E_CAPACITY; }
    if (n > 0) ORBX_CUDA(h, cudaMemcpy(out, db->d_out, (size_t)n * sizeof(orbx_dmatch), cudaMemcpyDeviceToHost));
    std::sort(out, out + n, [](const orbx_dmatch &a, const orbx_dmatch &b) { return a.queryIdx != b.queryIdx ? a.queryIdx < b.queryIdx : a.trainIdx < b.trainIdx; });
    return ORBX_OK;
}

// ---- reprojection-gated association (Backend::associateObservation, backend.cpp:1064-1120) ----
static orbx_status db_positions(orbx_db *db, int64_t first, int64_t n, const float *src, cudaMemcpyKind kind)
{
    if (!db || first < 0 || n < 0 || (n > 0 && !src)) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    if (first + n > db->cap) { h->err = "positions beyond the database capacity"; return ORBX_E_CAPACITY; }
    if (!db->d_pos) {
        ORBX_CUDA(h, cudaMalloc(&db->d_pos, (size_t)db->cap * 3 * sizeof(float)));
        ORBX_CUDA(h, cudaMemsetAsync(db->d_pos, 0, (size_t)db->cap * 3 * sizeof(float), h->stream));
    }
    if (n > 0) ORBX_CUDA(h, cudaMemcpyAsync(db->d_pos + (size_t)first * 3, src, (size_t)n * 3 * sizeof(float), kind, h->stream));
    if (kind == cudaMemcpyHostToDevice) ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}
extern "C" orbx_status orbx_db_set_positions(orbx_db *db, int64_t first, int64_t n, const float *xyz) { return db_positions(db, first, n, xyz, cudaMemcpyHostToDevice); }
extern "C" orbx_status orbx_db_set_positions_device(orbx_db *db, int64_t first, int64_t n, const float *d_xyz) { return db_positions(db, first, n, d_xyz, cudaMemcpyDeviceToDevice); }

extern "C" orbx_status orbx_db_associate_device(orbx_db *db, const uint8_t *d_q, const float *d_qpx, int32_t nq, const orbx_pose *pose,
                                                float max_dist, double max_err, orbx_assoc *d_out)
{
    if (!db || nq < 0 || !pose || !d_out || (nq > 0 && (!d_q || !d_qpx))) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    if (nq == 0) return ORBX_OK;
    if (db->rows > 0 && !db->d_pos) { h->err = "orbx_db_set_positions has not been called"; return ORBX_E_INVALID; }
    if (db->rows > 0x7FFFFFFF) { h->err = "shard too large"; return ORBX_E_UNSUPPORTED; }
    if (launch_assoc(h, d_q, d_qpx, nq, db->d_rows, db->d_pos, (int)db->rows, db->first_index, pose, max_dist, max_err, d_out) != 0) {
        h->err = "out of device memory (association scratch)"; return ORBX_E_NOMEM;
    }
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
extern "C" orbx_status orbx_db_associate(orbx_db *db, const uint8_t *q, const float *qpx, int32_t nq, const orbx_pose *pose,
                                         float max_dist, double max_err, orbx_assoc *out)
{
    if (!db || nq < 0 || !pose || !out || (nq > 0 && (!q || !qpx))) return ORBX_E_INVALID;
    orbx_handle *h = db->h; cudaSetDevice(h->device);
    if (nq == 0) return ORBX_OK;
    orbx_status st = db_stage_queries(db, q, nq, (size_t)nq * sizeof(orbx_assoc));
    if (st != ORBX_OK) return st;
    if ((size_t)nq * 2 * sizeof(float) > db->qpx_cap) {
        if (db->d_qpx) { cudaStreamSynchronize(h->stream); cudaFree(db->d_qpx); db->d_qpx = nullptr; db->qpx_cap = 0; }
        ORBX_CUDA(h, cudaMalloc(&db->d_qpx, (size_t)nq * 2 * sizeof(float))); db->qpx_cap = (size_t)nq * 2 * sizeof(float);
    }
    ORBX_CUDA(h, cudaMemcpyAsync(db->d_qpx, qpx, (size_t)nq * 2 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if ((st = orbx_db_associate_device(db, db->d_q, db->d_qpx, nq, pose, max_dist, max_err, (orbx_assoc *)db->d_out)) != ORBX_OK) return st;
    ORBX_CUDA(h, cudaMemcpyAsync(out, db->d_out, (size_t)nq * sizeof(orbx_assoc), cudaMemcpyDeviceToHost, h->stream));
    { const orbx_status ds = check_device_status(h); if (ds != ORBX_OK) return ds; }      // synchronises; a broken tensor-memory pipeline (never seen) must not return silently
    return ORBX_OK;
}
extern "C" orbx_status orbx_merge_assoc_device(orbx_handle *h, const orbx_assoc *d_parts, int32_t nshards, int32_t nq, orbx_assoc *d_out)
{
    if (!h || !d_parts || !d_out || nshards < 1 || nq < 0) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    launch_assoc_merge(h, d_parts, nshards, nq, d_out);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

// ---- keyframe packing (Frontend::publishKeyframe, frontend.cpp:731-776) ----
extern "C" orbx_status orbx_pack_keyframe_device(orbx_handle *h, int32_t nframes, const orbx_keypoint *d_kps, const uint8_t *d_desc,
                                                 const int32_t *d_counts, int32_t cap, const uint16_t *d_depth, int32_t width, int32_t height,
                                                 size_t dstep, size_t dfstride, const orbx_kfparams *K,
                                                 orbx_kfrecord *d_out, int32_t *d_nout, int32_t out_cap)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (nframes < 1 || !d_kps || !d_desc || !d_counts || cap < 1 || !d_depth || width < 1 || height < 1 || dstep < (size_t)width * 2 || (dstep & 1) ||
        !K || !d_out || !d_nout || out_cap < 1) { h->err = "bad keyframe packing arguments"; return ORBX_E_INVALID; }
    launch_pack_keyframe(h, nframes, d_kps, d_desc, d_counts, cap, d_depth, dstep, dfstride, width, height, K, d_out, d_nout, out_cap);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
extern "C" orbx_status orbx_pack_keyframe(orbx_handle *h, const orbx_keypoint *kps, const uint8_t *desc, int32_t n, const uint16_t *depth,
                                          int32_t width, int32_t height, size_t dstep, const orbx_kfparams *K,
                                          orbx_kfrecord *out, int32_t cap, int32_t *n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n_out) *n_out = 0;
    if (n < 0 || (n > 0 && (!kps || !desc)) || !depth || width < 1 || height < 1 || dstep < (size_t)width * 2 || !K || !out || !n_out || cap < 0) { h->err = "bad keyframe packing arguments"; return ORBX_E_INVALID; }
    if (n == 0) return ORBX_OK;
    if (n > h->max_kp) { h->err = "more keypoints than the handle's max_keypoints"; return ORBX_E_CAPACITY; }
    if (width > h->prm.max_width || height > h->prm.max_height) { h->err = "depth image larger than max_width x max_height"; return ORBX_E_INVALID; }
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    orbx_status st;
    const size_t dpitch = align_up((size_t)width * 2, 128);
    const size_t need = (size_t)h->max_kp * sizeof(orbx_kfrecord) + 64;
    if ((st = grow(h, (uint8_t **)&h->d_mout, &h->mout_cap, need)) != ORBX_OK) return st;
    orbx_kfrecord *d_rec = (orbx_kfrecord *)h->d_mout;
    const uint16_t *dd = (const uint16_t *)mapped_device_view(depth);           // pinned depth is gathered in place
    size_t dds = dstep;
    if (!dd) {
        ORBX_CUDA(h, cudaMemcpy2DAsync(h->d_depth_in, dpitch, depth, dstep, (size_t)width * 2, (size_t)height, cudaMemcpyHostToDevice, h->stream));
        dd = h->d_depth_in; dds = dpitch;
    }
    ORBX_CUDA(h, cudaMemcpyAsync(h->d_kps_all, kps, (size_t)n * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(h->d_desc_all, desc, (size_t)n * ORBX_DESC_BYTES, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(h->d_count_all, &n, sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    launch_pack_keyframe(h, 1, h->d_kps_all, h->d_desc_all, h->d_count_all, h->max_kp, dd, dds, 0, width, height, K, d_rec, h->d_mcount, h->max_kp);
    int32_t m = 0;
    ORBX_CUDA(h, cudaMemcpyAsync(&m, h->d_mcount, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    *n_out = m;
    if (m > cap) { h->err = "keyframe record capacity too small"; return ORBX_E_CAPACITY; }
    if (m > 0) ORBX_CUDA(h, cudaMemcpy(out, d_rec, (size_t)m * sizeof(orbx_kfrecord), cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

// ---- feature culling for the backend (frontend.cpp:1168-1218) ----
#define ORBX_CULL_MAX 8192
extern "C" orbx_status orbx_cull_keyframe_device(orbx_handle *h, const orbx_keypoint *d_kps, const uint8_t *d_desc, int32_t n,
                                                 const int32_t *d_match_query, int32_t n_matches, int32_t max_new, float min_response,
                                                 orbx_keypoint *d_out_kps, uint8_t *d_out_desc, int32_t *d_out_index, int32_t cap, int32_t *d_n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n < 0 || n_matches < 0 || max_new < 0 || (n > 0 && (!d_kps || !d_desc)) || (n_matches > 0 && !d_match_query) || !d_out_kps || !d_out_desc ||
        !d_n_out || cap < 0 || ((uintptr_t)d_desc & 3) || ((uintptr_t)d_out_desc & 3)) { h->err = "bad culling arguments"; return ORBX_E_INVALID; }
    if (n > ORBX_CULL_MAX) { h->err = "more than 8192 keypoints in one culling call"; return ORBX_E_CAPACITY; }
    if (n == 0 && n_matches > 0) { h->err = "matches given for an empty keypoint set"; return ORBX_E_INVALID; }
    if (n == 0) { ORBX_CUDA(h, cudaMemsetAsync(d_n_out, 0, sizeof(int32_t), h->stream)); return ORBX_OK; }
    if (launch_cull(h, d_kps, d_desc, n, d_match_query, n_matches, max_new, min_response, d_out_kps, d_out_desc, d_out_index, cap, d_n_out) != 0) {
        h->err = "shared-memory opt-in for the culling kernel failed"; return ORBX_E_CUDA;
    }
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
extern "C" orbx_status orbx_cull_keyframe(orbx_handle *h, const orbx_keypoint *kps, const uint8_t *desc, int32_t n,
                                          const int32_t *match_query, int32_t n_matches, int32_t max_new, float min_response,
                                          orbx_keypoint *out_kps, uint8_t *out_desc, int32_t *out_index, int32_t cap, int32_t *n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n_out) *n_out = 0;
    if (n < 0 || n_matches < 0 || max_new < 0 || (n > 0 && (!kps || !desc)) || (n_matches > 0 && !match_query) || !out_kps || !out_desc || !n_out || cap < 0) {
        h->err = "bad culling arguments"; return ORBX_E_INVALID;
    }
    if (n > ORBX_CULL_MAX) { h->err = "more than 8192 keypoints in one culling call"; return ORBX_E_CAPACITY; }
    for (int i = 0; i < n_matches; i++)
        if (match_query[i] < 0 || match_query[i] >= n) { h->err = "a match query index lies outside the keypoint array"; return ORBX_E_INVALID; }
    if (n == 0) return ORBX_OK;
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    orbx_status st;
    const int ocap = n_matches + std::min(max_new, n);
    const size_t in_k = align_up((size_t)n * sizeof(orbx_keypoint), 16), in_d = (size_t)n * ORBX_DESC_BYTES, in_q = align_up((size_t)std::max(n_matches, 1) * 4, 16);
    const size_t out_k = align_up((size_t)std::max(ocap, 1) * sizeof(orbx_keypoint), 16), out_d = (size_t)std::max(ocap, 1) * ORBX_DESC_BYTES, out_i = align_up((size_t)std::max(ocap, 1) * 4, 16);
    if ((st = grow(h, &h->d_mq, &h->mq_cap, in_k + in_d + in_q)) != ORBX_OK) return st;
    if ((st = grow(h, &h->d_mt, &h->mt_cap, out_k + out_d + out_i + 16)) != ORBX_OK) return st;
    orbx_keypoint *dk = (orbx_keypoint *)h->d_mq; uint8_t *dd = h->d_mq + in_k; int32_t *dq = (int32_t *)(h->d_mq + in_k + in_d);
    orbx_keypoint *ok = (orbx_keypoint *)h->d_mt; uint8_t *od = h->d_mt + out_k; int32_t *oi = (int32_t *)(h->d_mt + out_k + out_d); int32_t *on = (int32_t *)(h->d_mt + out_k + out_d + out_i);
    ORBX_CUDA(h, cudaMemcpyAsync(dk, kps, (size_t)n * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(dd, desc, (size_t)n * ORBX_DESC_BYTES, cudaMemcpyHostToDevice, h->stream));
    if (n_matches > 0) ORBX_CUDA(h, cudaMemcpyAsync(dq, match_query, (size_t)n_matches * 4, cudaMemcpyHostToDevice, h->stream));
    if (launch_cull(h, dk, dd, n, dq, n_matches, max_new, min_response, ok, od, oi, ocap, on) != 0) { h->err = "shared-memory opt-in for the culling kernel failed"; return ORBX_E_CUDA; }
    int32_t m = 0;
    ORBX_CUDA(h, cudaMemcpyAsync(&m, on, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    *n_out = m;
    if (m > cap) { h->err = "culled feature capacity too small"; return ORBX_E_CAPACITY; }
    if (m > 0) {
        ORBX_CUDA(h, cudaMemcpyAsync(out_kps, ok, (size_t)m * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(h, cudaMemcpyAsync(out_desc, od, (size_t)m * ORBX_DESC_BYTES, cudaMemcpyDeviceToHost, h->stream));
        if (out_index) ORBX_CUDA(h, cudaMemcpyAsync(out_index, oi, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return ORBX_OK;
}

// ---- geometric validation: hypothesis scoring (frontend.cpp:1134-1154) ----
extern "C" orbx_status orbx_fmat_score(orbx_handle *h, const float *pts1, const float *pts2, int32_t n, const double *F, int32_t nh, double threshold,
                                       int32_t *inlier_counts, int32_t *best, uint8_t *best_mask)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n < 0 || nh < 1 || !F || !best || (n > 0 && (!pts1 || !pts2)) || !(threshold >= 0)) { h->err = "bad fundamental-matrix scoring arguments"; return ORBX_E_INVALID; }
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    orbx_status st;
    const size_t bp = align_up((size_t)std::max(n, 1) * 8, 16), bf = align_up((size_t)nh * 72, 16), bc = align_up((size_t)nh * 4, 16), bm = align_up((size_t)nh * std::max(n, 1), 16);
    if ((st = grow(h, &h->d_mq, &h->mq_cap, 2 * bp + bf + bc + 16 + align_up((size_t)std::max(n, 1), 16))) != ORBX_OK) return st;
    if ((st = grow(h, &h->d_mt, &h->mt_cap, bm)) != ORBX_OK) return st;
    float *d_p1 = (float *)h->d_mq, *d_p2 = (float *)(h->d_mq + bp);
    double *d_F = (double *)(h->d_mq + 2 * bp);
    int32_t *d_c = (int32_t *)(h->d_mq + 2 * bp + bf), *d_b = (int32_t *)(h->d_mq + 2 * bp + bf + bc);
    uint8_t *d_bm = h->d_mq + 2 * bp + bf + bc + 16;
    if (n > 0) {
        ORBX_CUDA(h, cudaMemcpyAsync(d_p1, pts1, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
        ORBX_CUDA(h, cudaMemcpyAsync(d_p2, pts2, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    }
    ORBX_CUDA(h, cudaMemcpyAsync(d_F, F, (size_t)nh * 72, cudaMemcpyHostToDevice, h->stream));
    launch_fmat_score(h, d_p1, d_p2, n, d_F, nh, (float)(threshold * threshold), d_c, h->d_mt, d_b, d_bm);
    ORBX_CUDA(h, cudaMemcpyAsync(best, d_b, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (inlier_counts) ORBX_CUDA(h, cudaMemcpyAsync(inlier_counts, d_c, (size_t)nh * 4, cudaMemcpyDeviceToHost, h->stream));
    if (best_mask && n > 0) ORBX_CUDA(h, cudaMemcpyAsync(best_mask, d_bm, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

// cv::solvePnPRansac's scoring loop (frontend.cpp:906-923): nh poses (R row-major 3x3, t) against n 3D-2D correspondences
extern "C" orbx_status orbx_pnp_score(orbx_handle *h, const float *pts3d, const float *pts2d, int32_t n, const double *Rt, int32_t nh,
                                      double fx, double fy, double cx, double cy, double threshold, int32_t *inlier_counts, int32_t *best, uint8_t *best_mask)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n < 0 || nh < 1 || !Rt || !best || (n > 0 && (!pts3d || !pts2d)) || !(threshold >= 0)) { h->err = "bad pose scoring arguments"; return ORBX_E_INVALID; }
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    orbx_status st;
    const size_t b3 = align_up((size_t)std::max(n, 1) * 12, 16), b2 = align_up((size_t)std::max(n, 1) * 8, 16), bf = align_up((size_t)nh * 96, 16), bc = align_up((size_t)nh * 4, 16),
                 bm = align_up((size_t)nh * std::max(n, 1), 16);
    if ((st = grow(h, &h->d_mq, &h->mq_cap, b3 + b2 + bf + bc + 16 + align_up((size_t)std::max(n, 1), 16))) != ORBX_OK) return st;
    if ((st = grow(h, &h->d_mt, &h->mt_cap, bm)) != ORBX_OK) return st;
    float *d_p3 = (float *)h->d_mq, *d_p2 = (float *)(h->d_mq + b3);
    double *d_Rt = (double *)(h->d_mq + b3 + b2);
    int32_t *d_c = (int32_t *)(h->d_mq + b3 + b2 + bf), *d_b = (int32_t *)(h->d_mq + b3 + b2 + bf + bc);
    uint8_t *d_bm = h->d_mq + b3 + b2 + bf + bc + 16;
    if (n > 0) {
        ORBX_CUDA(h, cudaMemcpyAsync(d_p3, pts3d, (size_t)n * 12, cudaMemcpyHostToDevice, h->stream));
        ORBX_CUDA(h, cudaMemcpyAsync(d_p2, pts2d, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    }
    ORBX_CUDA(h, cudaMemcpyAsync(d_Rt, Rt, (size_t)nh * 96, cudaMemcpyHostToDevice, h->stream));
    launch_pnp_score(h, d_p3, d_p2, n, d_Rt, nh, fx, fy, cx, cy, (float)(threshold * threshold), d_c, h->d_mt, d_b, d_bm);
    ORBX_CUDA(h, cudaMemcpyAsync(best, d_b, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (inlier_counts) ORBX_CUDA(h, cudaMemcpyAsync(inlier_counts, d_c, (size_t)nh * 4, cudaMemcpyDeviceToHost, h->stream));
    if (best_mask && n > 0) ORBX_CUDA(h, cudaMemcpyAsync(best_mask, d_bm, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

// the 3D-2D correspondences of Frontend::estimateCameraPose (frontend.cpp:858-892), in match order
extern "C" orbx_status orbx_pnp_points(orbx_handle *h, const orbx_keypoint *prev_kps, int32_t n_prev, const orbx_keypoint *curr_kps, int32_t n_curr,
                                       const orbx_dmatch *matches, int32_t nm, const uint16_t *prev_depth, int32_t width, int32_t height, size_t dstep,
                                       float fx, float fy, float cx, float cy, float *pts3d, float *pts2d, int32_t *n_out)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (nm < 0 || n_prev < 0 || n_curr < 0 || !n_out || (nm > 0 && (!matches || !prev_kps || !curr_kps || !prev_depth || !pts3d || !pts2d)) || width < 1 || height < 1 ||
        dstep < (size_t)width * 2 || (dstep & 1)) { h->err = "bad correspondence arguments"; return ORBX_E_INVALID; }
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    *n_out = 0;
    if (nm == 0) return ORBX_OK;
    orbx_status st;
    const size_t bk0 = align_up((size_t)std::max(n_prev, 1) * sizeof(orbx_keypoint), 16), bk1 = align_up((size_t)std::max(n_curr, 1) * sizeof(orbx_keypoint), 16),
                 bmm = align_up((size_t)nm * sizeof(orbx_dmatch), 16), b3 = align_up((size_t)nm * 12, 16), b2 = align_up((size_t)nm * 8, 16);
    if ((st = grow(h, &h->d_mq, &h->mq_cap, bk0 + bk1 + bmm + b3 + b2 + 16)) != ORBX_OK) return st;
    if ((st = grow(h, (uint8_t **)&h->d_depth_in, &h->depth_cap, (size_t)height * dstep)) != ORBX_OK) return st;
    orbx_keypoint *d_k0 = (orbx_keypoint *)h->d_mq, *d_k1 = (orbx_keypoint *)(h->d_mq + bk0);
    orbx_dmatch *d_m = (orbx_dmatch *)(h->d_mq + bk0 + bk1);
    float *d_p3 = (float *)(h->d_mq + bk0 + bk1 + bmm), *d_p2 = (float *)(h->d_mq + bk0 + bk1 + bmm + b3);
    int32_t *d_n = (int32_t *)(h->d_mq + bk0 + bk1 + bmm + b3 + b2);
    ORBX_CUDA(h, cudaMemcpyAsync(d_k0, prev_kps, (size_t)n_prev * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(d_k1, curr_kps, (size_t)n_curr * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(d_m, matches, (size_t)nm * sizeof(orbx_dmatch), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(h->d_depth_in, prev_depth, (size_t)height * dstep, cudaMemcpyHostToDevice, h->stream));
    launch_pnp_points(h, d_k0, n_prev, d_k1, n_curr, d_m, nm, h->d_depth_in, width, height, dstep, fx, fy, cx, cy, d_p3, d_p2, d_n);
    int32_t n = 0;
    ORBX_CUDA(h, cudaMemcpyAsync(&n, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    st = check_device_status(h);                                   // synchronises; a match index outside its keypoint array -> ORBX_E_INVALID
    if (st != ORBX_OK) return st;
    if (n > 0) {
        ORBX_CUDA(h, cudaMemcpy(pts3d, d_p3, (size_t)n * 12, cudaMemcpyDeviceToHost));
        ORBX_CUDA(h, cudaMemcpy(pts2d, d_p2, (size_t)n * 8, cudaMemcpyDeviceToHost));
    }
    *n_out = n;
    return ORBX_OK;
}

extern "C" orbx_status orbx_fmat_ransac(orbx_handle *h, const float *pts1, const float *pts2, int32_t n, int32_t nh, double threshold, uint32_t seed,
                                        double *F_out, uint8_t *best_mask, int32_t *n_inliers)
{
    if (!h) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    if (n < 8 || nh < 1 || !pts1 || !pts2 || !F_out || !best_mask || !(threshold >= 0)) { h->err = "bad fundamental-matrix RANSAC arguments (n >= 8, nh >= 1)"; return ORBX_E_INVALID; }
    if (h->pending[0].active || h->pending[1].active) { h->err = "an asynchronous batch is outstanding: call orbx_batch_wait first"; return ORBX_E_INVALID; }
    orbx_status st;
    const size_t bp = align_up((size_t)n * 8, 16), bf = align_up((size_t)nh * 72, 16), bc = align_up((size_t)nh * 4, 16), bm = align_up((size_t)nh * n, 16);
    if ((st = grow(h, &h->d_mq, &h->mq_cap, 2 * bp + bf + bc + 16 + align_up((size_t)n, 16))) != ORBX_OK) return st;
    if ((st = grow(h, &h->d_mt, &h->mt_cap, bm)) != ORBX_OK) return st;
    float *d_p1 = (float *)h->d_mq, *d_p2 = (float *)(h->d_mq + bp);
    double *d_F = (double *)(h->d_mq + 2 * bp);
    int32_t *d_c = (int32_t *)(h->d_mq + 2 * bp + bf), *d_b = (int32_t *)(h->d_mq + 2 * bp + bf + bc);
    uint8_t *d_bm = h->d_mq + 2 * bp + bf + bc + 16;
    ORBX_CUDA(h, cudaMemcpyAsync(d_p1, pts1, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(d_p2, pts2, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    launch_fmat_hypotheses(h, d_p1, d_p2, n, nh, seed, d_F);
    launch_fmat_score(h, d_p1, d_p2, n, d_F, nh, (float)(threshold * threshold), d_c, h->d_mt, d_b, d_bm);
    int32_t best = 0;
    ORBX_CUDA(h, cudaMemcpyAsync(&best, d_b, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaMemcpyAsync(best_mask, d_bm, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBX_CUDA(h, cudaMemcpy(F_out, d_F + (size_t)best * 9, 72, cudaMemcpyDeviceToHost));
    if (n_inliers) ORBX_CUDA(h, cudaMemcpy(n_inliers, d_c + best, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

// ---- synthetic inputs ----
extern "C" orbx_status orbx_synth_gray_device(orbx_handle *h, uint32_t seed, int32_t first, int32_t n, int32_t w, int32_t hgt, uint8_t *d, size_t step, size_t fstride)
{
    if (!h || !d || n < 1 || w < 1 || hgt < 1 || step < (size_t)w) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    launch_synth_gray(h, seed, first, n, w, hgt, d, step, fstride);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
extern "C" orbx_status orbx_synth_depth_device(orbx_handle *h, uint32_t seed, int32_t first, int32_t n, int32_t w, int32_t hgt, uint16_t *d, size_t step, size_t fstride)
{
    if (!h || !d || n < 1 || w < 1 || hgt < 1 || step < (size_t)w * 2) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    launch_synth_depth(h, seed, first, n, w, hgt, d, step, fstride);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
extern "C" orbx_status orbx_synth_descriptors_device(orbx_handle *h, uint32_t seed, uint64_t first_row, int64_t nrows, uint8_t *d)
{
    if (!h || !d || nrows < 1) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    launch_synth_desc(h, seed, first_row, nrows, d);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

// ---- utilities ----
extern "C" void *orbx_alloc_pinned(size_t bytes) { void *p = nullptr; return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr; }
extern "C" void orbx_free_pinned(void *p) { if (p) cudaFreeHost(p); }
extern "C" void *orbx_alloc_device(orbx_handle *h, size_t bytes) { if (!h) return nullptr; cudaSetDevice(h->device); void *p = nullptr; return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr; }
extern "C" void orbx_free_device(orbx_handle *h, void *p) { if (h && p) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); cudaFree(p); } }
extern "C" orbx_status orbx_copy_to_device(orbx_handle *h, void *d, const void *s, size_t bytes)
{
    if (!h || !d || !s) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    ORBX_CUDA(h, cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}
extern "C" orbx_status orbx_copy_to_host(orbx_handle *h, void *d, const void *s, size_t bytes)
{
    if (!h || !d || !s) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    ORBX_CUDA(h, cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" orbx_status orbx_test_trig(orbx_handle *h, const float *in, int32_t n, float *oc, float *os)
{
    if (!h || !in || !oc || !os || n < 1) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    float *d = nullptr;
    ORBX_CUDA(h, cudaMalloc(&d, sizeof(float) * 3 * (size_t)n));
    cudaMemcpyAsync(d, in, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream);
    launch_test_trig(h, d, n, d + n, d + 2 * (size_t)n);
    cudaMemcpyAsync(oc, d + n, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(os, d + 2 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    ORBX_CUDA(h, e);
    return ORBX_OK;
}
extern "C" orbx_status orbx_test_atan2(orbx_handle *h, const float *y, const float *x, int32_t n, float *out)
{
    if (!h || !y || !x || !out || n < 1) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    float *d = nullptr;
    ORBX_CUDA(h, cudaMalloc(&d, sizeof(float) * 3 * (size_t)n));
    cudaMemcpyAsync(d, y, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d + n, x, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream);
    launch_test_atan2(h, d, d + n, n, d + 2 * (size_t)n);
    cudaMemcpyAsync(out, d + 2 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    ORBX_CUDA(h, e);
    return ORBX_OK;
}
extern "C" orbx_status orbx_test_trig_checksum(orbx_handle *h, uint32_t first, uint32_t last, uint64_t *sc, uint64_t *ss)
{
    if (!h || !sc || !ss || last < first) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    unsigned long long *d = nullptr, hs[2] = { 0, 0 };
    ORBX_CUDA(h, cudaMalloc(&d, 16));
    cudaMemsetAsync(d, 0, 16, h->stream);
    launch_trig_checksum(h, first, last, d);
    cudaMemcpyAsync(hs, d, 16, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    ORBX_CUDA(h, e);
    *sc = hs[0]; *ss = hs[1];
    return ORBX_OK;
}

// the distribution stage alone, on a caller-supplied candidate list (order-independent by construction)
extern "C" orbx_status orbx_test_quadtree(orbx_handle *h, const int32_t *xys, int32_t n, int32_t box_w, int32_t box_h,
                                          int32_t wcell, int32_t hcell, int32_t ncols, int32_t N, int32_t *out_xys, int32_t cap, int32_t *n_out)
{
    if (!h || !xys || !out_xys || !n_out || n < 0 || box_w < 1 || box_h < 1 || wcell < 1 || hcell < 1 || ncols < 1 || N < 0) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    FrameGeom G; memset(&G, 0, sizeof(G));
    G.nlevels = 1;
    LevelGeom &g = G.lv[0];
    g.w = box_w + 2 * ORBX_BORDER; g.h = box_h + 2 * ORBX_BORDER;
    g.wcell = wcell; g.hcell = hcell; g.ncols = ncols; g.N = N;
    g.nini = (int)roundf((float)box_w / box_h);
    if (g.nini < 1 || g.nini > 64) { h->err = "unsupported aspect"; return ORBX_E_UNSUPPORTED; }
    g.hx = (float)box_w / g.nini;
    g.cand_cap = (int)std::min<size_t>(h->cand_cap, 1u << 30); g.cand_off = 0;
    g.sel_cap = (int)align_up((size_t)std::max(N + 4, 4 * g.nini + 4), 8); g.sel_off = 0;
    if (n > g.cand_cap || (size_t)g.sel_cap > h->sel_cap || n > 65535 * 64) { h->err = "test input too large"; return ORBX_E_CAPACITY; }
    std::vector<uint32_t> packed((size_t)std::max(n, 1));
    for (int i = 0; i < n; i++) packed[i] = orbx_pack(xys[3 * i], xys[3 * i + 1], xys[3 * i + 2]);
    FrameGeom *d_g = nullptr;
    ORBX_CUDA(h, cudaMalloc(&d_g, sizeof(FrameGeom)));
    cudaMemcpyAsync(d_g, &G, sizeof(G), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_cand, packed.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_ncand, &n, sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
    if (launch_quadtree_geo(h, d_g, 1, 1, g.sel_cap, h->cand_cap, (int)h->sel_cap) != 0) { cudaFree(d_g); h->err = "quadtree node table exceeds the shared-memory opt-in limit"; return ORBX_E_UNSUPPORTED; }
    int32_t m = 0;
    cudaMemcpyAsync(&m, h->d_nsel, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(d_g);
    ORBX_CUDA(h, e);
    orbx_status st = check_device_status(h);
    if (st != ORBX_OK) return st;
    *n_out = m;
    if (m > cap) return ORBX_E_CAPACITY;
    std::vector<uint32_t> sel((size_t)std::max(m, 1));
    ORBX_CUDA(h, cudaMemcpy(sel.data(), h->d_sel, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) { out_xys[3 * i] = orbx_px(sel[i]); out_xys[3 * i + 1] = orbx_py(sel[i]); out_xys[3 * i + 2] = orbx_ps(sel[i]); }
    return ORBX_OK;
}

// ---- per-kernel event profiling ----
static const char *k_prof_names[ORBX_K_COUNT] = { "k_resize_linear", "k_fast_cells", "k_quadtree", "k_blur7", "k_describe", "k_filter",
                                                  "k_match_partial", "k_match_epilogue", "other", "k_fast_dense", "k_fast_nms", "k_fast_retry" };
extern "C" int32_t orbx_profile_kernels(void) { return ORBX_K_COUNT; }
extern "C" const char *orbx_profile_name(int32_t id) { return id >= 0 && id < ORBX_K_COUNT ? k_prof_names[id] : ""; }
extern "C" void orbx_profile_enable(orbx_handle *h, int32_t on)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (on && h->prof_ev.empty()) {
        h->prof_ev.resize(2 * 8192); h->prof_id.resize(8192);
        for (auto &e : h->prof_ev) cudaEventCreate(&e);
    }
    h->prof_on = on ? 1 : 0;
}
extern "C" orbx_status orbx_profile_read(orbx_handle *h, double *ms, int64_t *launches)
{
    if (!h || !ms || !launches) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < h->prof_n; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]) == cudaSuccess) { h->prof_ms[h->prof_id[i]] += t; h->prof_cnt[h->prof_id[i]]++; }
    }
    h->prof_n = 0;
    for (int k = 0; k < ORBX_K_COUNT; k++) { ms[k] = h->prof_ms[k]; launches[k] = h->prof_cnt[k]; h->prof_ms[k] = 0; h->prof_cnt[k] = 0; }
    return ORBX_OK;
}

extern "C" orbx_status orbx_bench_popc(orbx_handle *h, double *popc_per_sec)
{
    if (!h || !popc_per_sec) return ORBX_E_INVALID;
    cudaSetDevice(h->device);
    *popc_per_sec = run_popc_bench(h);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}
