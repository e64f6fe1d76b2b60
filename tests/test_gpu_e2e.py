"""GPU: the whole front end (Voice -> PQMF -> bridge -> VICReg loss) through the public API against the CPU oracle."""
import types

import numpy as np
import pytest
import torch

import harness

pytestmark = pytest.mark.gpu


def test_front_end_step_vs_oracle(cuda_device):
    import ias_b200

    B = 64
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=True, sample_rate=44100, buffer_size_seconds=4.0)
    voice = ias_b200.Voice(synthconfig=cfg).to(cuda_device)
    gram = ias_b200.PQMF(N=3).to(cuda_device)
    vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(cuda_device)
    audio, params, is_train = voice(0)
    bands = gram(audio.unsqueeze(1))
    assert bands.shape == (B, 3, 58800)
    x, y = harness.bridge(bands, params, wa, wp)
    # the bench's step pools inside the analysis kernel: same bands bit for bit, same embeddings to fp32 rounding
    bands_f, x_f, y_f = harness.analysis_bridge(gram, audio, params, wa, wp)
    assert torch.equal(bands_f, bands) and torch.equal(y_f, y)
    assert float((x_f - x).abs().max() / x.abs().max()) <= 1e-5
    with torch.no_grad():
        out = vic.loss(x, y)
        out_f = vic.loss(x_f, y_f)
    got = np.array([float(o) for o in out])
    assert np.all(np.abs(np.array([float(o) for o in out_f]) - got) <= 1e-4 * np.abs(got))
    ref = harness.oracle_front_end(0, B, N=3)
    want = np.array(ref["loss4"])
    rel = np.abs(got - want) / np.abs(want)
    print("front end loss terms", got, want, rel)
    # bands: the few voices whose audio differs (ill-conditioned pitch path) move the pooled features slightly
    band_err = float((bands.cpu() - ref["bands"]).abs().max() / ref["bands"].abs().max())
    print("bands rel err", band_err, "embedding rel err",
          float((x.cpu() - ref["x"]).abs().max() / ref["x"].abs().max()))
    assert np.all(rel <= 2e-5)  # measured 1e-6 (loss terms average the few ill-conditioned voices away); bound 1e-4
    # loss on the ORACLE embeddings isolates the loss kernels: north-star tolerance
    with torch.no_grad():
        out2 = vic.loss(ref["x"].to(cuda_device), ref["y"].to(cuda_device))
    got2 = np.array([float(o) for o in out2])
    assert np.all(np.abs(got2 - want) <= 1e-4 * np.abs(want))
    # PQMF on the ORACLE audio isolates the filter kernel
    bands2 = gram(ref["audio"].unsqueeze(1).to(cuda_device))
    assert float((bands2.cpu() - ref["bands"]).abs().max() / ref["bands"].abs().max()) <= 1e-5
    # the image view the reference's AudioEmbedding takes (audioembed.py:41) is a pure reshape of the bands
    assert bands.reshape(-1, 3, 240, 245).shape == (B, 3, 240, 245)


def test_smoke_entry(cuda_device):
    import __graft_entry__ as g

    g.smoke()


def test_bridge_pool_kernel_matches_torch(cuda_device):
    """The harness pooling kernel == torch's adaptive_avg_pool1d(|x|) (bins overlap by one element when S % P != 0)."""
    for B, N, L, P in [(8, 3, 58800, 256), (3, 16, 11025, 256), (2, 1, 1000, 7), (2, 2, 256, 256)]:
        g = torch.Generator().manual_seed(L)
        bands = (torch.rand((B, N, L), generator=g) * 2 - 1)
        ref = torch.nn.functional.adaptive_avg_pool1d(bands.abs().reshape(B, 1, -1).double(), P).squeeze(1)
        wa = torch.eye(P, device=cuda_device)
        harness_dim = harness.EMBED_DIM
        try:
            harness.EMBED_DIM = P
            x, _ = harness.bridge(bands.to(cuda_device), torch.zeros(B, 78, device=cuda_device), wa,
                                  torch.zeros(78, P, device=cuda_device))
        finally:
            harness.EMBED_DIM = harness_dim
        assert float((x.cpu().double() - ref).abs().max() / ref.abs().max()) <= 1e-6


def test_device_batch_index_and_cuda_graph_replay(cuda_device):
    """Voice(batch_idx) with the batch number resident on the device (ias_voice_seed_params_dev): same result as the
    host integer, and the whole step replays from a CUDA graph with a new batch number each time."""
    import ias_b200

    B = 32
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=True, sample_rate=44100, buffer_size_seconds=1.0)
    voice = ias_b200.Voice(synthconfig=cfg).to(cuda_device)
    gram = ias_b200.PQMF(N=3).to(cuda_device)
    vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(cuda_device)

    def step(idx):
        audio, params, is_train = voice(idx)
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp)
        with torch.no_grad():
            return audio, params, is_train, torch.stack(vic.loss(x, y))

    want = {i: [t.clone() for t in step(i)] for i in (3, 11, 2 ** 33 + 5)}
    idx_dev = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    for i, w in want.items():  # eager, device-resident index
        idx_dev.fill_(i)
        got = step(idx_dev)
        for a, b in zip(got, w):
            assert torch.equal(a, b)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(idx_dev)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static = step(idx_dev)
    host_idx = torch.zeros(1, dtype=torch.int64).pin_memory()
    for i, w in want.items():
        host_idx[0] = i
        idx_dev.copy_(host_idx, non_blocking=True)
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip(static, w):
            assert torch.equal(a, b)


def test_second_device_in_the_same_process(cuda_device):
    """One process driving two GPUs: calls made while cuda:0 is the current device must still run on the GPU that
    holds the tensors (the bindings switch device around every library call; opt-in shared memory is set per device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ias_b200

    B = 32
    outs = []
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        cfg = ias_b200.SynthConfig(batch_size=B, reproducible=True, sample_rate=44100, buffer_size_seconds=4.0)
        voice = ias_b200.Voice(synthconfig=cfg).to(dev)
        gram = ias_b200.PQMF(N=3).to(dev)
        vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
            mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
        vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity(), gather=False)
        wa, wp = harness.bridge_weights(dev)
        assert torch.cuda.current_device() == 0
        audio, params, _ = voice(2)
        bands = gram(audio.unsqueeze(1))
        x, y = harness.bridge(bands, params, wa, wp)
        x.requires_grad_(True)
        loss = vic.loss(x, y)
        loss[0].backward()
        assert audio.device == dev and bands.device == dev and loss[0].device == dev
        outs.append((audio.cpu(), bands.cpu(), torch.stack([l.detach() for l in loss]).cpu(), x.grad.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_prepared_batches_render_bit_identically(cuda_device):
    """Voice.prepare(k+1) on a side stream while batch k is consumed, then Voice(k+1, prepared=True): the audio stage
    alone renders exactly what Voice(k+1) renders (audio, parameters, is_train, loss), eagerly and from a CUDA graph."""
    import ias_b200

    B = 32
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=True, sample_rate=44100, buffer_size_seconds=4.0)
    voice = ias_b200.Voice(synthconfig=cfg).to(cuda_device)
    gram = ias_b200.PQMF(N=3).to(cuda_device)
    vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(cuda_device)

    def tail(audio, params):
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp)
        with torch.no_grad():
            return torch.stack(vic.loss(x, y))

    batches = [4, 5, 6, 7]
    want = []
    for k in batches:
        audio, params, is_train = voice(k)
        want.append((audio.clone(), params.clone(), is_train.clone(), tail(audio, params).clone()))

    side = torch.cuda.Stream(device=cuda_device)
    voice.prepare(batches[0])
    for n, k in enumerate(batches):
        audio, params, is_train = voice(k, prepared=True)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            voice.prepare(k + 1)
        loss = tail(audio, params)
        cur.wait_stream(side)
        for got, ref in zip((audio, params, is_train, loss), want[n]):
            assert torch.equal(got, ref)
    # a batch that was not prepared is refused
    voice.prepare(3)
    with pytest.raises(RuntimeError):
        voice(9, prepared=True)
    audio3, _, _ = voice(3, prepared=True)  # the refused call left the prepared state alone
    assert torch.equal(audio3, voice(3)[0])
    with pytest.raises(RuntimeError):
        voice(3, prepared=True)  # consumed

    # the same pipeline captured in a CUDA graph with a device-resident batch number
    idx = torch.zeros(1, dtype=torch.int64, device=cuda_device)

    def step():
        audio, params, is_train = voice(idx, prepared=True)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            idx.add_(1)
            voice.prepare(idx)
        loss = tail(audio, params)
        cur.wait_stream(side)
        return audio, params, is_train, loss

    def prime():
        idx.fill_(batches[0])
        voice.prepare(idx)

    prime()
    warm = torch.cuda.Stream(device=cuda_device)
    warm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(warm):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(warm)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    prime()
    for n in range(len(batches)):
        g.replay()
        torch.cuda.synchronize()
        for got, ref in zip(out, want[n]):
            assert torch.equal(got, ref)


def test_two_deep_pipeline_renders_bit_identically(cuda_device):
    """The control stage in two halves, two batches ahead (bench.py --pipeline-depth 2): modulation stage of batch k+1
    under the analysis of batch k, seeding + ADSR envelopes of batch k+2 after it.  Audio, parameters, is_train and the
    loss of every batch equal the unpipelined render bit for bit, eagerly and from a CUDA graph; the deferred
    normalisation factors ride along."""
    import ias_b200

    B = 32
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=True, sample_rate=44100, buffer_size_seconds=4.0)
    voice = ias_b200.Voice(synthconfig=cfg, normalize="defer").to(cuda_device)
    gram = ias_b200.PQMF(N=3).to(cuda_device)
    vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(cuda_device)
    side = torch.cuda.Stream(device=cuda_device)

    def tail(audio, params, hook=None):
        bands, x, y = harness.analysis_bridge(gram, audio, params, wa, wp, voice.row_scale, after_analysis=hook)
        with torch.no_grad():
            return torch.stack(vic.loss(x, y))

    batches = [4, 5, 6, 7, 8]
    want = []
    for k in batches:
        audio, params, is_train = voice(k)
        want.append((audio.clone(), params.clone(), is_train.clone(), tail(audio, params).clone()))

    with pytest.raises(RuntimeError):
        voice.prepare_modulation()  # nothing enveloped yet
    idx = torch.zeros(1, dtype=torch.int64, device=cuda_device)

    def prime():
        idx.fill_(batches[0])
        voice.prepare_envelopes(idx)
        voice.prepare_modulation()
        idx.add_(1)
        voice.prepare_envelopes(idx)

    def step():
        audio, params, is_train = voice(idx, prepared=True)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            voice.prepare_modulation()
        done = torch.cuda.Event()
        loss = tail(audio, params, hook=lambda: done.record(cur))
        with torch.cuda.stream(side):
            side.wait_event(done)
            idx.add_(1)
            voice.prepare_envelopes(idx)
        cur.wait_stream(side)
        return audio, params, is_train, loss

    prime()
    for n in range(len(batches)):
        out = step()
        torch.cuda.synchronize()
        for got, ref in zip(out, want[n]):
            assert torch.equal(got, ref)

    prime()
    warm = torch.cuda.Stream(device=cuda_device)
    warm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(warm):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(warm)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    prime()
    for n in range(len(batches)):
        g.replay()
        torch.cuda.synchronize()
        for got, ref in zip(out, want[n]):
            assert torch.equal(got, ref)
