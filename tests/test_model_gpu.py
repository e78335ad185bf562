"""Device-resident forward (llmi_model_*) against the reference's Model::forward.

Bars: greedy argmax tokens identical along the decode (north_star: first 32
steps).  Logits: the typical step agrees to ~1e-7 of max|logit|; what cannot be
bounded tightly is the reference's own fp16 arithmetic — K/V are stored as f16
and the attention value accumulator is rounded to f16 at every cached position
(model.cpp:461-474, 528-538) — where a 1e-7 upstream difference (summation
order of the mat-vecs, libm vs CUDA tanhf/expf/sincosf) occasionally flips one
rounding (2^-11 relative on that element, persistent once it sits in the KV
cache; SURVEY §7 hard part 7).  So: MEDIAN error <= 1e-5, every step <= 5e-2 —
and tests/test_reference_conditioning.py shows that the reference
itself moves by the same order under a 1-ulp change of one weight, i.e. the 5e-2
bar is the conditioning of the reference's fp16 arithmetic, not slack for ours."""
import numpy as np
import pytest

from llm_inference_b200 import synth

pytestmark = pytest.mark.gpu
REPO_GOLDEN = __import__("pathlib").Path(__file__).resolve().parent / "golden" / "model_golden.npz"
TOL_TYPICAL, TOL_WORST = 1e-5, 5e-2


@pytest.fixture(scope="module")
def mg():
    return np.load(REPO_GOLDEN)


@pytest.mark.parametrize("name", ["q4_0", "q4_k_m"])
def test_forward_matches_reference_golden(gpu_ops, mg, name):
    from llm_inference_b200.model import Model
    m = Model(mg[f"{name}_image"], max_positions=64)
    ref_logits, ref_tokens, prompt = mg[f"{name}_logits"], mg[f"{name}_tokens"], mg[f"{name}_prompt"]
    lg = m.forward(prompt, 0)
    errs = [float(np.abs(lg - ref_logits[0]).max() / np.abs(ref_logits[0]).max())]
    pos = len(prompt)
    for i, t_ref in enumerate(ref_tokens):
        t = int(lg.argmax())
        assert t == int(t_ref), f"greedy token {i} differs"
        lg = m.forward([t], pos)
        errs.append(float(np.abs(lg - ref_logits[i + 1]).max() / np.abs(ref_logits[i + 1]).max()))
        pos += 1
    assert max(errs) <= TOL_WORST and float(np.median(errs)) <= TOL_TYPICAL, errs
    m.close()


@pytest.mark.parametrize("name", ["q4_0", "q4_k_m"])
def test_device_greedy_loop_equals_host_loop(gpu_ops, mg, name):
    from llm_inference_b200.model import Model
    m = Model(mg[f"{name}_image"], max_positions=64)
    prompt, ref_tokens = mg[f"{name}_prompt"], mg[f"{name}_tokens"]
    lg = m.forward(prompt, 0)
    first = int(lg.argmax())
    toks, ms = m.decode_greedy(first, len(prompt), len(ref_tokens) - 1)
    assert first == int(ref_tokens[0])
    assert list(toks) == [int(t) for t in ref_tokens[1:]]
    assert ms > 0 and m.launches_per_step > 0
    # the graph is reusable: run again from the same state
    m.forward(prompt, 0)
    toks2, _ = m.decode_greedy(first, len(prompt), len(ref_tokens) - 1)
    assert np.array_equal(toks, toks2)
    last = m.last_logits()
    # the last executed step consumed ref_tokens[-2]: its logits are golden row len-1
    assert np.abs(last - mg[f"{name}_logits"][len(ref_tokens) - 1]).max() <= TOL_WORST * float(np.abs(last).max())
    m.close()


def test_32_greedy_steps_token_identical_vs_compiled_reference(gpu_ops):
    """north_star: greedy argmax tokens must match over the first 32 decode steps.
    Two regimes: embeddings ~N(0,1) as SURVEY §8(d) specifies (the tied logits then
    mostly echo the input token), and small embeddings (the network decides)."""
    from oracle import binding
    if not binding.ref_available():
        pytest.skip("oracle/_ref did not travel to this box")
    from llm_inference_b200.model import Model
    dims = synth.GemmaDims("small", 3, 512, 1024, 4, 2, 128, 512)
    R = binding.Ref(n_threads=8)
    for wt, et, seed, std in ((synth.Q4_0, synth.F16, 1, 1.0), (synth.Q4_0, synth.F16, 2, 0.004),
                              ("q4_k_m", synth.Q6_K, 3, 0.004), (synth.Q8_0, synth.Q8_0, 4, 0.004)):
        img = synth.build_gemma3_gguf(dims, wt, et, seed=seed, embd_std=std)
        ref, m = R.model(img), Model(img, max_positions=128)
        prompt = np.arange(5, 21, dtype=np.int32)
        a, b = ref.forward(prompt, 0), m.forward(prompt, 0)
        pos, margins, errs, toks, near_ties = len(prompt), [], [], [], 0
        for step in range(32):
            ta, tb = int(a.argmax()), int(b.argmax())
            srt = np.sort(a)
            margins.append(float((srt[-1] - srt[-2]) / np.abs(a).max()))
            errs.append(float(np.abs(a - b).max() / np.abs(a).max()))
            near_ties += margins[-1] <= 2 * errs[-1]  # top-1 / top-2 closer than the observed error: reported, not excused
            assert ta == tb, f"{wt}: token {step} differs (margin {margins[-1]:.3e}, err {errs[-1]:.3e})"
            toks.append(ta)
            a, b = ref.forward([ta], pos), m.forward([ta], pos)
            pos += 1
        # Nothing is excused: the tokens above are asserted at EVERY step.  near_ties only reports how many steps were
        # decided by less than twice the logit error — on the small-embedding models that error sits at the reference's
        # own chaos floor (tests/test_reference_conditioning.py: ~2e-2 under a 1-ulp perturbation, because an int8
        # activation rounding or an f16 KV rounding that flips is a 2^-8 .. 2^-11 relative step), so such steps exist.
        # N(0,1) embeddings: logits are dominated by well-conditioned terms; small embeddings make the
        # logits tiny next to the activations, so one flipped f16 rounding in the KV cache shows as ~1e-2
        assert max(errs) <= (1e-3 if std == 1.0 else TOL_WORST), errs
        print(f"{wt} std={std}: 32 greedy tokens identical ({len(set(toks))} distinct), min margin/max "
              f"{min(margins):.2e}, logits err/max: median {np.median(errs):.1e} max {max(errs):.1e}, steps with margin <= "
              f"2 x err: {near_ties}")
        ref.close()
        m.close()


def test_model_load_errors(gpu_ops):
    from llm_inference_b200 import _lib
    from llm_inference_b200.model import Model
    g = synth.GGUFBuilder()
    g.add_str("general.architecture", "gemma4")
    with pytest.raises(_lib.LlmiError, match="only the gemma3 architecture"):
        Model(g.build())
    g = synth.GGUFBuilder()
    g.add_str("general.architecture", "gemma3")
    with pytest.raises(_lib.LlmiError, match="Failed to find metadata key: gemma3.block_count"):
        Model(g.build())
    with pytest.raises(_lib.LlmiError, match="Invalid GGUF magic number"):
        Model(np.zeros(64, np.uint8))


@pytest.mark.parametrize("wt,et", [(synth.Q4_0, synth.F16), ("q4_k_m", synth.Q6_K), (synth.Q8_0, synth.Q8_0)])
def test_batched_prefill_is_bitwise_the_token_by_token_path(gpu_ops, monkeypatch, wt, et):
    """A prompt goes through each layer in batches (weights read once per token tile); per token the
    arithmetic is that of the one-token path, so logits AND the KV cache (probed by decoding on) must be
    bit-identical, for a batch size that splits the prompt unevenly."""
    from llm_inference_b200.model import Model
    dims = synth.GemmaDims("small", 3, 512, 1024, 4, 2, 256, 640)
    img = synth.build_gemma3_gguf(dims, wt, et, seed=11, embd_std=0.02)
    prompt = (np.arange(37, dtype=np.int32) * 7 + 3) % dims.vocab
    outs = {}
    # batch 16: token-per-lane dp4a kernel; batch 36: the tcgen05 int8 kernel (>= 32 tokens) on a ragged token tile
    for mode, env in (("batched", {"LLMI_PREFILL_BATCH": "16"}), ("umma", {"LLMI_PREFILL_BATCH": "36", "LLMI_UMMA_MIN_TOKENS": "32"}),
                      ("single", {"LLMI_NO_PREFILL": "1"})):
        for k in ("LLMI_PREFILL_BATCH", "LLMI_NO_PREFILL", "LLMI_UMMA_MIN_TOKENS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = Model(img, max_positions=64)
        lg = [m.forward(prompt, 0)]
        ms, launches = m.last_forward_stats()
        pos = len(prompt)
        for _ in range(3):
            lg.append(m.forward([int(lg[-1].argmax())], pos))
            pos += 1
        lg.append(m.forward(prompt[:9], pos))  # a second, short prompt appended to the same cache
        outs[mode] = (np.stack(lg), launches)
        m.close()
    a, b = outs["batched"][0], outs["single"][0]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(outs["umma"][0].view(np.uint32), b.view(np.uint32))
    assert outs["batched"][1] < outs["single"][1] / 4  # 3 batches of launches instead of 37 tokens' worth


@pytest.mark.parametrize("dims", [synth.GemmaDims("small", 3, 512, 1024, 4, 2, 128, 640),
                                  synth.GemmaDims("wide", 2, 2560, 1024, 4, 2, 256, 640)], ids=lambda d: d.name)
def test_norm_stage_fused_into_the_ring_kernel_is_bitwise_the_two_kernels(gpu_ops, monkeypatch, dims):
    """llmi_launch_gemv_batch_norm: the RMSNorm (+ residual + next norm + Q8_0 quantizer) stage as the prologue of the
    persistent mat-vec launch it feeds, every CTA computing its own copy of the activation in shared memory, the
    residual stream alternating between two buffers.  Forced onto every launch of a small model (ring mode 2; the ring
    kernel is by default only taken by >= 10 MB Q4_0 launches, and the fusion is opt-in: LLMI_FUSE_NORM=1, measured slower)
    and compared BITWISE with the separate kernels (ring mode 1, and
    LLMI_FUSE_NORM=0 with the ring on): logits token by token and the greedy decode after them; 512 logical norm threads
    (E = 512) and 1024 (E = 2560: two logical threads per physical thread)."""
    from llm_inference_b200.model import Model
    img = synth.build_gemma3_gguf(dims, synth.Q4_0, synth.F16, seed=41)
    prompt = (np.arange(9, dtype=np.int32) * 13 + 1) % dims.vocab
    outs = {}
    monkeypatch.setenv("LLMI_NO_PREFILL", "1")  # the prompt goes token by token through run_step
    for mode, ring, fuse in (("fused", 2, "1"), ("ring_unfused", 2, "0"), ("slab", 1, "1")):
        monkeypatch.setenv("LLMI_FUSE_NORM", fuse)
        gpu_ops.set_gemv_ring(ring, 0, 0, 0)
        m = Model(img, max_positions=48)
        lg = [m.forward(prompt, 0)]
        pos = len(prompt)
        for _ in range(3):
            lg.append(m.forward([int(lg[-1].argmax())], pos))
            pos += 1
        toks, _ = m.decode_greedy(int(lg[-1].argmax()), pos, 6)
        outs[mode] = (np.stack(lg), toks, m.launches_per_step)
        m.close()
    gpu_ops.set_gemv_ring(0, 0, 0, 0)
    monkeypatch.delenv("LLMI_FUSE_NORM", raising=False)
    for other in ("ring_unfused", "slab"):
        assert np.array_equal(outs["fused"][0].view(np.uint32), outs[other][0].view(np.uint32)), other
        assert np.array_equal(outs["fused"][1], outs[other][1]), other
    assert outs["fused"][2] < outs["slab"][2]  # two launches fewer per layer


@pytest.mark.parametrize("wt,et", [(synth.Q4_0, synth.F16), ("q4_k_m", synth.Q6_K)])
def test_cluster_norm_is_bitwise_the_single_cta_norm(gpu_ops, monkeypatch, wt, et):
    """norm_act_cluster_kernel (8 CTAs, distributed shared memory; used from E = 2048 on) spreads the single-CTA kernel's
    1024 logical threads over a cluster with the same per-thread sums, shuffle trees and left-to-right warp sums: logits
    of a prompt (batched: one cluster per token) and of the decode steps after it must be bit-identical, for the Q8_0
    (registers) and the Q8_K (through the fp32 copy) activation emitters."""
    from llm_inference_b200.model import Model
    dims = synth.GemmaDims("wide", 2, 2560, 1024, 4, 2, 256, 640)
    img = synth.build_gemma3_gguf(dims, wt, et, seed=31)
    prompt = (np.arange(19, dtype=np.int32) * 11 + 5) % dims.vocab
    outs = {}
    for mode in ("cluster", "single"):
        monkeypatch.delenv("LLMI_NORM_CLUSTER", raising=False)
        if mode == "single":
            monkeypatch.setenv("LLMI_NORM_CLUSTER", "0")
        m = Model(img, max_positions=48)
        lg = [m.forward(prompt, 0)]
        pos = len(prompt)
        for _ in range(3):
            lg.append(m.forward([int(lg[-1].argmax())], pos))
            pos += 1
        toks, _ = m.decode_greedy(int(lg[-1].argmax()), pos, 6)
        outs[mode] = (np.stack(lg), toks)
        m.close()
    monkeypatch.delenv("LLMI_NORM_CLUSTER", raising=False)
    assert np.array_equal(outs["cluster"][0].view(np.uint32), outs["single"][0].view(np.uint32))
    assert np.array_equal(outs["cluster"][1], outs["single"][1])


@pytest.mark.parametrize("wt,et,hd", [(synth.Q4_0, synth.F16, 128), ("q4_k_m", synth.Q6_K, 256), (synth.Q8_0, synth.Q8_0, 64)])
def test_throughput_prefill_mode_stays_within_its_tolerance(gpu_ops, monkeypatch, wt, et, hd):
    """LLMI_PREFILL=fast (opt-in): prompts of >= 64 tokens go through the dequantize-to-bf16 tcgen05 GEMM and the fp32
    online-softmax attention.  Not the parity path: the bar it states is |logits - exact| <= 2e-2 * max|exact| after a
    150-token prompt (N(0,1) embeddings), the KV cache it leaves must carry the decode on (3 exact-path steps, same
    bar), and the greedy tokens are REPORTED against the exact path (equal on these seeds), not excused."""
    from llm_inference_b200.model import Model
    dims = synth.GemmaDims("small", 3, 512, 1024, 4, 2, hd, 640)
    img = synth.build_gemma3_gguf(dims, wt, et, seed=23)
    prompt = (np.arange(150, dtype=np.int32) * 7 + 3) % dims.vocab
    outs = {}
    for mode in ("exact", "fast"):
        monkeypatch.delenv("LLMI_PREFILL", raising=False)
        if mode == "fast":
            monkeypatch.setenv("LLMI_PREFILL", "fast")
        m = Model(img, max_positions=192)
        lg = [m.forward(prompt, 0)]
        _, launches = m.last_forward_stats()
        pos = len(prompt)
        for _ in range(3):
            lg.append(m.forward([int(lg[-1].argmax())], pos))
            pos += 1
        outs[mode] = (np.stack(lg), launches)
        m.close()
    monkeypatch.delenv("LLMI_PREFILL", raising=False)
    gpu_ops.set_prefill_mode(False)
    e, f = outs["exact"][0], outs["fast"][0]
    assert not np.isnan(f).any()
    errs = [float(np.abs(f[i] - e[i]).max() / np.abs(e[i]).max()) for i in range(len(e))]
    assert max(errs) <= 2e-2, errs
    assert not np.array_equal(e[0], f[0])  # another arithmetic: equality would mean the exact path ran
    same = [int(e[i].argmax()) == int(f[i].argmax()) for i in range(len(e))]
    print(f"{wt}: fast vs exact logits err/max {['%.1e' % v for v in errs]}, greedy tokens equal {same}, launches {outs['fast'][1]} vs {outs['exact'][1]}")
    assert all(same)


@pytest.mark.parametrize("wt,et,hd", [(synth.Q4_0, synth.F16, 128), ("q4_k_m", synth.Q6_K, 256), (synth.Q8_0, synth.Q8_0, 64),
                                      (synth.Q5_0, synth.Q5_0, 128)])
def test_persistent_decode_kernel_is_bitwise_the_per_launch_path(gpu_ops, monkeypatch, wt, et, hd):
    """The persistent decode kernel (mega.cu: one launch per decode call, flagged dataflow between the stages,
    norms as mat-vec prologues, GEGLU as the gate/up epilogue) must give the bits of the per-launch path
    (model.cu run_step: one kernel per stage) — logits, greedy tokens and KV cache — for any number of CTAs."""
    from llm_inference_b200.model import Model
    dims = synth.GemmaDims("small", 3, 512, 1024, 4, 2, hd, 648)
    img = synth.build_gemma3_gguf(dims, wt, et, seed=21, embd_std=0.02)
    prompt = (np.arange(11, dtype=np.int32) * 5 + 2) % dims.vocab
    outs = {}
    for mode, env in (("legacy", {"LLMI_DECODE": "legacy"}), ("mega", {"LLMI_DECODE": "mega"}),
                      ("mega_5ctas", {"LLMI_DECODE": "mega", "LLMI_MEGA_CTAS": "5"}),
                      ("mega_prompt", {"LLMI_DECODE": "mega", "LLMI_NO_PREFILL": "1"})):
        for k in ("LLMI_DECODE", "LLMI_MEGA_CTAS", "LLMI_NO_PREFILL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = Model(img, max_positions=64)
        lg0 = m.forward(prompt, 0)              # legacy batched prefill, or (mega_prompt) token by token in the kernel
        toks, _ = m.decode_greedy(int(lg0.argmax()), len(prompt), 12)
        last = m.last_logits()
        pos = len(prompt) + 12
        lgs = [m.forward([int(toks[-1])], pos)]  # the reference-facing call on top of the cache the kernel wrote
        lgs.append(m.forward([int(lgs[-1].argmax())], pos + 1))
        outs[mode] = (lg0, toks, last, np.stack(lgs), m.launches_per_step)
        m.close()
    ref = outs["legacy"]
    assert ref[4] > 8  # the per-launch path really ran
    for mode in ("mega", "mega_5ctas", "mega_prompt"):
        o = outs[mode]
        assert o[4] == 1, mode
        assert np.array_equal(o[0].view(np.uint32), ref[0].view(np.uint32)), f"{mode}: prompt logits"
        assert np.array_equal(o[1], ref[1]), f"{mode}: greedy tokens {o[1]} vs {ref[1]}"
        assert np.array_equal(o[2].view(np.uint32), ref[2].view(np.uint32)), f"{mode}: last logits of the greedy loop"
        assert np.array_equal(o[3].view(np.uint32), ref[3].view(np.uint32)), f"{mode}: forward() logits after decode"


def _ref_greedy(ref_model, prompt, n_steps):
    """The reference's generation loop (main.cpp:165-221): the n_steps + 1 greedy tokens after the prompt, the
    top-1 / top-2 margin of every step (relative to max |logit|) and the logits of the last step."""
    lg = ref_model.forward(prompt, 0)
    pos, toks, margins = len(prompt), [], []
    for _ in range(n_steps + 1):
        srt = np.sort(lg)
        margins.append(float((srt[-1] - srt[-2]) / np.abs(lg).max()))
        toks.append(int(lg.argmax()))
        if len(toks) == n_steps + 1:
            break
        lg = ref_model.forward([toks[-1]], pos)
        pos += 1
    return toks, margins, lg


@pytest.mark.parametrize("workload,prompt_len,steps", [("gemma-3-1b-q4_0", 64, 32), ("gemma-3-4b-q4_k_m", 16, 8)])
def test_free_running_greedy_decode_at_baseline_dims_matches_reference(gpu_ops, workload, prompt_len, steps):
    """BASELINE.json configs 2 and 3 at FULL dimensions (gemma-3-1b: 26 layers, E 1152, F 6912, 4/1 heads x 256,
    V 262144, Q4_0 + F16 logits; gemma-3-4b Q4_K_M layout: 34 layers, E 2560, mixed Q4_K / Q6_K, Q6_K embeddings):
    prompt through the batched prefill, then the on-device greedy loop FREE-RUNNING (every token feeds the next step,
    nothing is teacher-forced) against the reference's own generation loop on its CPU code (oracle/_ref).  Bar
    (north_star): identical tokens; the minimum top-1 / top-2 margin is reported."""
    from oracle import binding
    if not binding.ref_available():
        pytest.skip("oracle/_ref did not travel to this box")
    import bench
    from llm_inference_b200.model import Model
    dims_name, wt, et = bench.WORKLOADS[workload]
    dims = synth.GEMMA3[dims_name]
    # embeddings small next to the layer outputs, so that the tied logits depend on the whole network instead of echoing
    # the input token (layers >= 1 reuse layer 0's generated bytes: generation time only)
    img = synth.build_gemma3_gguf(dims, wt, et, seed=4321, embd_std=0.02, distinct_layers=False)
    prompt = ((np.arange(prompt_len, dtype=np.int64) * 7919 + 13) % dims.vocab).astype(np.int32)
    R = binding.Ref(n_threads=__import__("os").cpu_count() or 1)
    ref = R.model(img)
    ref_toks, margins, ref_last = _ref_greedy(ref, prompt, steps)
    ref.close()
    m = Model(img, max_positions=prompt_len + steps + 8)
    first = int(m.forward(prompt, 0).argmax())
    toks, _ = m.decode_greedy(first, prompt_len, steps)
    ours = [first] + [int(t) for t in toks]
    last = m.last_logits()  # logits of the step that consumed ours[-2]: they produced ours[-1]
    m.close()
    assert ours == ref_toks, f"{workload}: free-running greedy tokens differ\nours {ours}\nref  {ref_toks}\nmargins {margins}"
    err = float(np.abs(last - ref_last).max() / np.abs(ref_last).max())
    assert err <= TOL_WORST, err
    print(f"{workload}: {steps + 1} free-running greedy tokens identical ({len(set(ours))} distinct), min margin / max|logit| "
          f"{min(margins):.2e}, last-step logits err / max {err:.1e}")
