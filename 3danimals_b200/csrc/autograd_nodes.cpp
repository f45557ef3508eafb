// autograd_nodes.cpp - the autograd layer of the hot path in C++ (torch::autograd::Function) over the C-ABI of libb2a.so.
//
// Why: the M1a step is host-bound: ~1.0 ms of Python / autograd / launch work per step against 0.6 ms of kernels.  A
// Python torch.autograd.Function costs ~10-15 us per node and direction in interpreter + engine work, every torch.empty from Python
// ~2.8 us; the four heaviest nodes of the step (fused render geometry, skinning, the antialias pair, marching tets) are ~330 us of
// it.  The same nodes here allocate through ATen and call the same entry points of include/b2a.h - no kernel lives in this file, the
// Python `Function`s of 3danimals_b200/ops.py remain the reference implementation (B2A_CPP_NODES=0 selects them; bench.py uses them
// for its per-call event timing).  PyTorch is plumbing: device memory, streams, the autograd tape.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>

#include <map>
#include <stdexcept>
#include <vector>

#include "../../include/b2a.h"

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

namespace {

int64_t g_launches = 0;                  // kernels launched through this module (bench.py's gpu_launches)
std::map<std::string, int64_t> g_calls;  // calls per entry point

void check(int rc)
{
    if (rc) throw std::runtime_error(std::string(b2a_last_error_string()));
}
b2a_stream_t stream() { return (b2a_stream_t)c10::cuda::getCurrentCUDAStream().stream(); }
void count(const char* name, int launches)
{
    g_launches += launches;
    g_calls[name] += 1;
}
const void* P(const Tensor& t) { return t.defined() ? t.data_ptr() : nullptr; }
void* PM(Tensor& t) { return t.defined() ? t.data_ptr() : nullptr; }

Tensor f32c(const Tensor& t, const char* name)
{
    if (!t.is_cuda()) throw std::runtime_error(std::string(name) + " must be a CUDA tensor (the B200 hot path has no CPU fallback)");
    if (t.scalar_type() == torch::kFloat32 && t.is_contiguous()) return t;
    return t.to(torch::kFloat32).contiguous();
}
Tensor empty_f32(at::IntArrayRef shape, const Tensor& like) { return torch::empty(shape, like.options().dtype(torch::kFloat32)); }
Tensor bytes(size_t n, const Tensor& like) { return torch::empty({(int64_t)std::max<size_t>(n, 16)}, like.options().dtype(torch::kUInt8)); }

// per-(device, stream) [B,V,12] vertex-gradient accumulator kept ZEROED between calls (the backward's per-vertex pass re-zeroes it)
std::map<std::pair<int, void*>, Tensor> g_acc;
Tensor gb_accumulator(size_t nbytes, const Tensor& like)
{
    auto key = std::make_pair((int)like.device().index(), (void*)stream());
    auto it = g_acc.find(key);
    if (it == g_acc.end() || (size_t)it->second.numel() < nbytes) {
        Tensor t = torch::zeros({(int64_t)std::max<size_t>(nbytes, 16)}, like.options().dtype(torch::kUInt8));
        g_acc[key] = t;
        return t;
    }
    return it->second;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Fused geometry half of render_mesh (ops._RenderGeometry): b2a_render_geometry_fwd / bwd
// outputs: clip, rast, aa_ctx (uint8, empty when no analysis), then the requested g-buffers in GB order (pos, geo, shn, cam, tex)
// ---------------------------------------------------------------------------------------------------------------------------
struct RenderGeometry : public torch::autograd::Function<RenderGeometry> {
    static variable_list forward(AutogradContext* ctx, Tensor v_pos, Tensor v_nrm, Tensor prior_pos, Tensor mtx, Tensor w2c, Tensor campos, Tensor tri,
                                 Tensor opp, int64_t H, int64_t W, int64_t spp, bool two_sided, int64_t want_mask, bool need_aa)
    {
        v_pos = f32c(v_pos, "v_pos"); v_nrm = f32c(v_nrm, "v_nrm"); prior_pos = f32c(prior_pos, "prior_pos");
        mtx = f32c(mtx, "matrix"); w2c = f32c(w2c, "w2c"); campos = f32c(campos, "campos");
        const int64_t B = mtx.size(0), V = v_pos.size(1), Bq = prior_pos.size(0), F = tri.size(0);
        if (v_pos.size(0) != B || w2c.size(0) != B || campos.size(0) != B || v_nrm.sizes() != v_pos.sizes())
            throw std::runtime_error("render geometry: inconsistent batch shapes");
        const int64_t fH = H * spp, fW = W * spp;
        size_t n;
        Tensor clip = empty_f32({B, V, 4}, v_pos);
        check(b2a_rasterize_workspace_bytes((int)B, F, (int)fH, (int)fW, &n));
        Tensor ws = bytes(n, v_pos);
        Tensor rast = empty_f32({B, fH, fW, 4}, v_pos);
        const bool use_cov = spp == 1;
        Tensor cov_list, cov_count;
        if (use_cov) {
            cov_list = torch::empty({B * fH * fW, 4}, v_pos.options().dtype(torch::kInt32));
            cov_count = torch::empty({1}, v_pos.options().dtype(torch::kInt32));
        }
        Tensor outs[5];
        for (int k = 0; k < 5; k++)
            if (want_mask & (1 << k)) outs[k] = empty_f32({B, H, W, 3}, v_pos);
        check(b2a_gbuffer_pack_bytes((int)B, (int)Bq, V, &n));
        Tensor packed = bytes(n, v_pos);
        Tensor aa_ctx;
        if (need_aa && (fH * fW) % 32 == 0 && F < (1ll << 28)) {
            check(b2a_antialias_workspace_bytes((int)B, (int)fH, (int)fW, &n));
            aa_ctx = bytes(n, v_pos);
        }
        check(b2a_render_geometry_fwd((const float*)P(v_pos), (const float*)P(v_nrm), (const float*)P(prior_pos), (int)Bq, (const float*)P(mtx),
                                      (const float*)P(w2c), (const float*)P(campos), (const int32_t*)P(tri), (const int32_t*)P(opp), two_sided ? 1 : 0, (int)B, V,
                                      F, (int)H, (int)W, (int)spp, PM(ws), ws.numel(), PM(packed), packed.numel(), (float*)PM(clip), (float*)PM(rast),
                                      (int32_t*)PM(cov_list), (int32_t*)PM(cov_count), (float*)PM(outs[0]), (float*)PM(outs[1]), (float*)PM(outs[2]),
                                      (float*)PM(outs[3]), (float*)PM(outs[4]), PM(aa_ctx), aa_ctx.defined() ? aa_ctx.numel() : 0, stream()));
        count("b2a_render_geometry_fwd", aa_ctx.defined() ? 8 : 6);
        ctx->save_for_backward({rast, clip, tri, v_pos, v_nrm, prior_pos, mtx, w2c, campos, use_cov ? cov_list : Tensor(), use_cov ? cov_count : Tensor(), packed});
        ctx->saved_data["spp"] = spp; ctx->saved_data["two_sided"] = two_sided; ctx->saved_data["H"] = H; ctx->saved_data["W"] = W;
        ctx->saved_data["want"] = want_mask;
        ctx->set_materialize_grads(false);
        Tensor aa_out = aa_ctx.defined() ? aa_ctx : torch::empty({0}, v_pos.options().dtype(torch::kUInt8));
        ctx->mark_non_differentiable({aa_out});
        variable_list res = {clip, rast, aa_out};
        for (int k = 0; k < 5; k++)
            if (outs[k].defined()) res.push_back(outs[k]);
        return res;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads)
    {
        auto sv = ctx->get_saved_variables();
        Tensor rast = sv[0], clip = sv[1], tri = sv[2], v_pos = sv[3], v_nrm = sv[4], prior_pos = sv[5], mtx = sv[6], w2c = sv[7], campos = sv[8],
               cov_list = sv[9], cov_count = sv[10], packed = sv[11];
        const int64_t spp = ctx->saved_data["spp"].toInt(), H = ctx->saved_data["H"].toInt(), W = ctx->saved_data["W"].toInt(),
                      want = ctx->saved_data["want"].toInt();
        const bool two_sided = ctx->saved_data["two_sided"].toBool();
        const int64_t B = v_pos.size(0), V = v_pos.size(1), F = tri.size(0);
        Tensor up = grads[0].defined() ? f32c(grads[0], "d_clip") : Tensor();
        Tensor d_rast = grads[1];
        Tensor gs[5];
        size_t gi = 3;
        bool have_gb = false;
        for (int k = 0; k < 5; k++)
            if (want & (1 << k)) {
                if (gi < grads.size() && grads[gi].defined()) { gs[k] = f32c(grads[gi], "d_gb"); have_gb = true; }
                gi++;
            }
        const bool need_nrm = ctx->needs_input_grad(1), need_prior = ctx->needs_input_grad(2), need_mtx = ctx->needs_input_grad(3),
                   need_w2c = ctx->needs_input_grad(4), need_cam = ctx->needs_input_grad(5);
        Tensor d_v_pos = torch::empty_like(v_pos);
        Tensor d_v_nrm = need_nrm ? torch::empty_like(v_nrm) : Tensor();
        Tensor d_prior = need_prior ? torch::empty_like(prior_pos) : Tensor();
        Tensor d_mtx = need_mtx ? torch::zeros_like(mtx) : Tensor();
        Tensor d_w2c = need_w2c ? torch::zeros_like(w2c) : Tensor();
        Tensor d_campos = need_cam ? torch::zeros_like(campos) : Tensor();
        if (d_rast.defined()) {     // only the 'flow' mode interpolates with a differentiable rast (render.py:281-288)
            Tensor d_clip = torch::zeros_like(clip);
            Tensor dr = f32c(d_rast, "d_rast");
            check(b2a_rasterize_bwd((const float*)P(clip), (const int32_t*)P(tri), (const float*)P(rast), (const float*)P(dr), (int)B, V, F, (int)rast.size(1),
                                    (int)rast.size(2), (float*)PM(d_clip), stream()));
            count("b2a_rasterize_bwd", 1);
            up = up.defined() ? up + d_clip : d_clip;
        }
        size_t n;
        check(b2a_gbuffer_bwd_workspace_bytes((int)B, V, &n));
        Tensor acc = gb_accumulator(n, v_pos);
        check(b2a_render_geometry_bwd((const float*)P(rast), (int)spp, (const float*)P(mtx), (const int32_t*)P(tri), (const float*)P(v_pos), (const float*)P(v_nrm),
                                      (const float*)P(prior_pos), (int)prior_pos.size(0), (const float*)P(w2c), (const float*)P(campos), two_sided ? 1 : 0, (int)B,
                                      V, F, (int)H, (int)W, P(packed), packed.numel(), (const int32_t*)P(cov_list), (const int32_t*)P(cov_count),
                                      (const float*)P(gs[0]), (const float*)P(gs[1]), (const float*)P(gs[2]), (const float*)P(gs[3]), (const float*)P(gs[4]),
                                      (const float*)P(up), PM(acc), acc.numel(), 1, (float*)PM(d_v_pos), (float*)PM(d_v_nrm), (float*)PM(d_prior),
                                      (float*)PM(d_mtx), (float*)PM(d_w2c), (float*)PM(d_campos), stream()));
        count("b2a_render_geometry_bwd", have_gb ? 2 : 1);
        return {ctx->needs_input_grad(0) ? d_v_pos : Tensor(), d_v_nrm, d_prior, d_mtx, d_w2c, d_campos, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(),
                Tensor(), Tensor(), Tensor()};
    }
};

// ---------------------------------------------------------------------------------------------------------------------------
// Linear blend skinning (ops._LBS): bone transforms + LBS, one node.  v_pos [Bv,V,3], bones [Bb,K,2,3], angles [B,K,3]
// -> out [B,V,3], posed [B,K,2,3]
// ---------------------------------------------------------------------------------------------------------------------------
struct LBS : public torch::autograd::Function<LBS> {
    static variable_list forward(AutogradContext* ctx, Tensor v_pos, Tensor bones, Tensor angles, Tensor chain_ptr, Tensor chain_ids, double temperature)
    {
        v_pos = f32c(v_pos, "v_pos"); bones = f32c(bones, "bones"); angles = f32c(angles, "angles");
        const int64_t B = angles.size(0), K = angles.size(1), Bv = v_pos.size(0), V = v_pos.size(1), Bb = bones.size(0);
        Tensor T_local = empty_f32({B, K, 12}, v_pos), G = empty_f32({B, K, 12}, v_pos), posed = empty_f32({B, K, 2, 3}, v_pos);
        check(b2a_lbs_bone_transforms((const float*)P(bones), (const float*)P(angles), (const int32_t*)P(chain_ptr), (const int32_t*)P(chain_ids), (int)B, (int)Bb,
                                      (int)K, (float*)PM(T_local), (float*)PM(G), (float*)PM(posed), stream()));
        count("b2a_lbs_bone_transforms", 2);
        Tensor out = empty_f32({B, V, 3}, v_pos);
        const float inv_t = (float)(1.0 / temperature);
        check(b2a_lbs_fwd((const float*)P(v_pos), (const float*)P(bones), (const float*)P(G), (int)B, (int)Bv, (int)Bb, (int)K, V, inv_t, (float*)PM(out), nullptr,
                          stream()));
        count("b2a_lbs_fwd", 1);
        ctx->save_for_backward({v_pos, bones, angles, chain_ptr, chain_ids, T_local, G});
        ctx->saved_data["inv_t"] = (double)inv_t;
        return {out, posed};
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads)
    {
        auto sv = ctx->get_saved_variables();
        Tensor v_pos = sv[0], bones = sv[1], angles = sv[2], chain_ptr = sv[3], chain_ids = sv[4], T_local = sv[5], G = sv[6];
        const float inv_t = (float)ctx->saved_data["inv_t"].toDouble();
        const int64_t B = angles.size(0), K = angles.size(1), Bv = v_pos.size(0), V = v_pos.size(1), Bb = bones.size(0);
        const bool need_v = ctx->needs_input_grad(0), need_a = ctx->needs_input_grad(2);
        const int64_t n = B * K * 12;
        Tensor zbuf = torch::zeros({2 * n + (need_v ? Bv * V * 3 : 0)}, v_pos.options());     // one zero fill for d_G, d_T, d_v
        Tensor d_G = zbuf.narrow(0, 0, n).view({B, K, 12});
        Tensor d_v = need_v ? zbuf.narrow(0, 2 * n, Bv * V * 3).view({Bv, V, 3}) : Tensor();
        if (grads[0].defined() && V > 0) {
            Tensor g = f32c(grads[0], "d_out");
            check(b2a_lbs_bwd((const float*)P(v_pos), (const float*)P(bones), (const float*)P(G), (const float*)P(g), (int)B, (int)Bv, (int)Bb, (int)K, V, inv_t,
                              (float*)PM(d_v), (float*)PM(d_G), stream()));
            count("b2a_lbs_bwd", 1);
        }
        Tensor d_angles;
        if (need_a) {
            Tensor d_T = zbuf.narrow(0, n, n).view({B, K, 12});
            d_angles = empty_f32({B, K, 3}, v_pos);
            Tensor gp = grads[1].defined() ? f32c(grads[1], "d_posed") : Tensor();
            check(b2a_lbs_bone_transforms_bwd((const float*)P(bones), (const float*)P(angles), (const int32_t*)P(chain_ptr), (const int32_t*)P(chain_ids),
                                              (const float*)P(T_local), (float*)PM(d_G), (const float*)P(gp), (int)B, (int)Bb, (int)K, (float*)PM(d_T),
                                              (float*)PM(d_angles), stream()));
            count("b2a_lbs_bone_transforms_bwd", 2);
        }
        return {d_v, Tensor(), d_angles, Tensor(), Tensor(), Tensor()};
    }
};

// ---------------------------------------------------------------------------------------------------------------------------
// composite + antialias of the training pair (ops._AntialiasPair): wide key (dino, 16+1) + narrow key (shaded, 3+1), NCHW views out
// ---------------------------------------------------------------------------------------------------------------------------
struct AntialiasPair : public torch::autograd::Function<AntialiasPair> {
    static variable_list forward(AutogradContext* ctx, Tensor color_w, Tensor color_n, c10::optional<Tensor> bg_w_, c10::optional<Tensor> bg_n_, Tensor pos,
                                 int64_t keep_w, int64_t keep_n, Tensor aa_ctx)
    {
        color_w = f32c(color_w, "color"); color_n = f32c(color_n, "color"); pos = f32c(pos, "pos");
        Tensor bg_w = bg_w_.has_value() && bg_w_->defined() ? f32c(*bg_w_, "background") : Tensor();
        Tensor bg_n = bg_n_.has_value() && bg_n_->defined() ? f32c(*bg_n_, "background") : Tensor();
        const int64_t B = color_w.size(0), H = color_w.size(1), W = color_w.size(2), Cw = color_w.size(3) + 1, Cn = color_n.size(3) + 1;
        const int Bgw = bg_w.defined() ? (int)bg_w.size(0) : 1, Bgn = bg_n.defined() ? (int)bg_n.size(0) : 1;
        Tensor out_w = empty_f32({B, H, W, Cw}, color_w), out_n = empty_f32({B, H, W, Cn}, color_w);
        check(b2a_antialias_pair_fwd((const float*)P(color_w), (const float*)P(bg_w), Bgw, (int)Cw, (float*)PM(out_w), (const float*)P(color_n), (const float*)P(bg_n),
                                     Bgn, (int)Cn, (float*)PM(out_n), (int)B, (int)H, (int)W, PM(aa_ctx), aa_ctx.numel(), stream()));
        count("b2a_antialias_pair_fwd", 1);
        ctx->save_for_backward({color_w, color_n, bg_w, bg_n, pos, aa_ctx});
        ctx->saved_data["keep_w"] = keep_w; ctx->saved_data["keep_n"] = keep_n;
        // (needs_input_grad indexes the VARIABLE inputs only, and the optional backgrounds may or may not be among them)
        ctx->saved_data["need_pos"] = pos.requires_grad();
        Tensor ow = keep_w < Cw ? out_w.narrow(3, 0, keep_w) : out_w, on = keep_n < Cn ? out_n.narrow(3, 0, keep_n) : out_n;
        return {ow.permute({0, 3, 1, 2}), on.permute({0, 3, 1, 2})};      // render.py:334 hands NCHW views of the NHWC storage to the caller
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads)
    {
        auto sv = ctx->get_saved_variables();
        Tensor color_w = sv[0], color_n = sv[1], bg_w = sv[2], bg_n = sv[3], pos = sv[4], aa_ctx = sv[5];
        const int64_t keep_w = ctx->saved_data["keep_w"].toInt(), keep_n = ctx->saved_data["keep_n"].toInt();
        const int64_t B = color_w.size(0), H = color_w.size(1), W = color_w.size(2), Cw = color_w.size(3) + 1, Cn = color_n.size(3) + 1, V = pos.size(1);
        const int Bgw = bg_w.defined() ? (int)bg_w.size(0) : 1, Bgn = bg_n.defined() ? (int)bg_n.size(0) : 1;
        Tensor g_w = grads[0].defined() ? grads[0] : torch::zeros({B, keep_w, H, W}, color_w.options());
        Tensor g_n = grads[1].defined() ? grads[1] : torch::zeros({B, keep_n, H, W}, color_w.options());
        if (g_w.scalar_type() != torch::kFloat32) g_w = g_w.to(torch::kFloat32);
        if (g_n.scalar_type() != torch::kFloat32) g_n = g_n.to(torch::kFloat32);
        auto fused_layout = [&](const Tensor& g) {      // g: [B,C,H,W]-shaped.  NCHW rows (x stride 1) or NHWC-contiguous storage
            const bool nchw_ok = g.stride(3) == 1 && W % 32 == 0;
            const bool nhwc_ok = g.stride(1) == 1 && g.stride(3) == keep_w && g.stride(2) == W * keep_w && g.stride(0) % 4 == 0 && ((uintptr_t)g.data_ptr()) % 16 == 0;
            return nchw_ok || nhwc_ok;
        };
        if (!fused_layout(g_w)) { g_w = g_w.contiguous(); g_n = g_n.contiguous(); }     // any other layout: one copy to NCHW rows (W % 32 == 0 here)
        g_w = g_w.permute({0, 2, 3, 1}); g_n = g_n.permute({0, 2, 3, 1});          // [B,H,W,C]-shaped views: the strides carry the layout
        const int64_t wsb = g_w.stride(0), wsy = g_w.stride(1), wsx = g_w.stride(2), wsc = g_w.stride(3);
        const int64_t nsb = g_n.stride(0), nsy = g_n.stride(1), nsx = g_n.stride(2), nsc = g_n.stride(3);
        const bool nhwc = wsc == 1 && wsx == keep_w && wsy == W * keep_w && wsb % 4 == 0 && ((uintptr_t)g_w.data_ptr()) % 16 == 0;
        const bool nchw = wsx == 1 && W % 32 == 0;
        if (!((nhwc || nchw) && keep_w == Cw - 1 && (keep_n == Cn || keep_n == Cn - 1)))
            throw std::runtime_error("antialias pair backward: gradient layout without a fused instantiation (set B2A_CPP_NODES=0)");
        Tensor d_color_w = torch::empty_like(color_w), d_color_n = torch::empty_like(color_n);
        Tensor d_pos = ctx->saved_data["need_pos"].toBool() ? torch::zeros_like(pos) : Tensor();
        check(b2a_antialias_pair_bwd((const float*)P(color_w), (const float*)P(bg_w), Bgw, (int)Cw, (const float*)P(g_w), wsb, wsy, wsx, wsc, (int)keep_w,
                                     (float*)PM(d_color_w), (const float*)P(color_n), (const float*)P(bg_n), Bgn, (int)Cn, (const float*)P(g_n), nsb, nsy, nsx, nsc,
                                     (int)keep_n, (float*)PM(d_color_n), (int)B, V, (int)H, (int)W, (float*)PM(d_pos), PM(aa_ctx), aa_ctx.numel(), stream()));
        count("b2a_antialias_pair_bwd", 1);
        return {d_color_w, d_color_n, Tensor(), Tensor(), d_pos, Tensor(), Tensor(), Tensor()};
    }
};

// ---------------------------------------------------------------------------------------------------------------------------
// Vertex normals (ops._VertexNormals)
// ---------------------------------------------------------------------------------------------------------------------------
struct VertexNormals : public torch::autograd::Function<VertexNormals> {
    static Tensor forward(AutogradContext* ctx, Tensor v_pos, Tensor tri)
    {
        v_pos = f32c(v_pos, "v_pos");
        const int64_t B = v_pos.size(0), V = v_pos.size(1), F = tri.size(0);
        Tensor nsum = empty_f32({B, V, 4}, v_pos), nrm = torch::empty_like(v_pos);
        check(b2a_vertex_normals_fwd((const float*)P(v_pos), (const int32_t*)P(tri), (int)B, V, F, (float*)PM(nsum), (float*)PM(nrm), stream()));
        count("b2a_vertex_normals_fwd", 2);
        ctx->save_for_backward({v_pos, tri, nsum});
        return nrm;
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads)
    {
        auto sv = ctx->get_saved_variables();
        Tensor v_pos = sv[0], tri = sv[1], nsum = sv[2];
        Tensor g = f32c(grads[0], "d_nrm");
        Tensor scratch = torch::empty_like(nsum), d_pos = torch::zeros_like(v_pos);
        check(b2a_vertex_normals_bwd((const float*)P(v_pos), (const int32_t*)P(tri), (const float*)P(nsum), (const float*)P(g), (int)v_pos.size(0), v_pos.size(1),
                                     tri.size(0), (float*)PM(scratch), (float*)PM(d_pos), stream()));
        count("b2a_vertex_normals_bwd", 2);
        return {d_pos, Tensor()};
    }
};

variable_list render_geometry(Tensor v_pos, Tensor v_nrm, Tensor prior_pos, Tensor mtx, Tensor w2c, Tensor campos, Tensor tri, Tensor opp, int64_t H, int64_t W,
                              int64_t spp, bool two_sided, int64_t want_mask, bool need_aa)
{
    return RenderGeometry::apply(v_pos, v_nrm, prior_pos, mtx, w2c, campos, tri, opp, H, W, spp, two_sided, want_mask, need_aa);
}
variable_list lbs(Tensor v_pos, Tensor bones, Tensor angles, Tensor chain_ptr, Tensor chain_ids, double temperature)
{
    return LBS::apply(v_pos, bones, angles, chain_ptr, chain_ids, temperature);
}
variable_list antialias_pair(Tensor color_w, Tensor color_n, c10::optional<Tensor> bg_w, c10::optional<Tensor> bg_n, Tensor pos, int64_t keep_w, int64_t keep_n,
                             Tensor aa_ctx)
{
    return AntialiasPair::apply(color_w, color_n, bg_w, bg_n, pos, keep_w, keep_n, aa_ctx);
}
Tensor vertex_normals(Tensor v_pos, Tensor tri) { return VertexNormals::apply(v_pos, tri); }

}  // namespace

PYBIND11_MODULE(_b2a_autograd, m)
{
    m.def("render_geometry", &render_geometry);
    m.def("lbs", &lbs);
    m.def("antialias_pair", &antialias_pair);
    m.def("vertex_normals", &vertex_normals);
    m.def("launches", []() { return g_launches; });
    m.def("calls", []() { return g_calls; });
    m.def("reset_stats", []() { g_launches = 0; g_calls.clear(); });
}
