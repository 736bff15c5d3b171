"""Host logic of the headless GUI audio path (streamvoiceanon_b200/realtime.py; reference evaluations/real-time-gui.py:32-49,
1204-1287, 1316-1359) with a stand-in model set -- no GPU.  The GPU run of the same path is
tests/test_zz_gpu_speaker.py::test_realtime_gui_glue."""
import numpy as np
import torch

from streamvoiceanon_b200 import realtime
from streamvoiceanon_b200.realtime import GuiState, RealtimeSession, custom_infer


class FakeModelSet:
    """The InferenceWrapper surface the GUI uses; process_one_chunk returns 2 * chunk + number of chunks since the last prompt."""
    sr = 44100
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []
        self.n = 0

    def prefill_prompt(self, ref, max_prompt_frames, delay, alpha):
        assert ref.dim() == 2 and ref.shape[0] == 1
        self.calls.append(("prompt", ref.shape[1], max_prompt_frames, delay, alpha))
        self.n = 0

    def setup_stream_caches(self, **kw):
        self.calls.append(("setup", kw))

    def process_one_chunk(self, wav):
        assert wav.dim() == 2 and wav.shape[0] == 1 and wav.shape[1] % 2048 == 0
        self.n += 1
        self.calls.append(("chunk", wav.shape[1]))
        return 2.0 * wav + self.n


def test_custom_infer_prompts_on_new_reference_or_block_size():
    ms, st = FakeModelSet(), GuiState()
    ref = np.zeros(3000, dtype=np.float32)
    out = custom_infer(ms, ref, "a.wav", torch.ones(2048), 2, 0.7, state=st)
    assert out.shape == (2048,) and float(out[0]) == 3.0
    assert ms.calls[0] == ("prompt", 3000, 64, 2, 0.7)                        # real-time-gui.py:37-40: max_prompt_frames=64
    assert ms.calls[1] == ("setup", dict(encode_window_frames=64, decode_window_frames=64, max_seq_frames=768, buffer_frames=32,
                                         decode_chunk_frames=1))                # :41-45
    custom_infer(ms, ref, "a.wav", torch.ones(2048), 2, 0.7, state=st)
    assert [c[0] for c in ms.calls] == ["prompt", "setup", "chunk", "chunk"]    # same reference and block size: no new prompt
    custom_infer(ms, ref, "b.wav", torch.ones(2048), 2, 0.7, state=st)         # another reference file
    custom_infer(ms, ref, "b.wav", torch.ones(4096), 2, 0.7, state=st)         # another block size
    kinds = [c[0] for c in ms.calls]
    assert kinds.count("prompt") == 3 and ms.calls[-2][1]["decode_chunk_frames"] == 2 and st.decode_chunk_frames == 2
    # without `state` the module-level instance is used, like the reference's globals
    realtime._STATE.reference_wav_name = ""
    custom_infer(ms, ref, "c.wav", torch.ones(2048))
    assert realtime._STATE.reference_wav_name == "c.wav"


def test_realtime_session_warm_up_blocks_and_callback():
    ms = FakeModelSet()
    sess = RealtimeSession(ms, samplerate=None, channels=2, block_frame=1, n_frame_delay=2, alpha=0.5)
    assert sess.samplerate == 44100 and sess.block_frame == 2048
    sess.start(np.zeros(5000, dtype=np.float32), "ref.wav")
    kinds = [c[0] for c in ms.calls]
    assert kinds == ["prompt", "setup"] + ["chunk"] * 5 and sess.resampler2 is None      # n_frame_delay + 3 silent blocks (:1263-1272)
    assert sess.state.reference_wav_name == ""                                             # the first real block prompts again
    indata = np.stack([np.full(2048, 0.25, np.float32), np.full(2048, 0.75, np.float32)], axis=1)    # [frames, channels]
    outdata = np.zeros((2048, 2), np.float32)
    sess.audio_callback(indata, outdata)
    assert [c[0] for c in ms.calls][-3:] == ["prompt", "setup", "chunk"]
    assert np.allclose(outdata, 2.0 * 0.5 + 1.0) and outdata.shape == (2048, 2)            # mono mix 0.5, both channels alike
    sess.audio_callback(indata, outdata)
    assert np.allclose(outdata, 2.0 * 0.5 + 2.0) and sess.infer_ms > 0.0
    sess.function = "passthrough"                                                          # any other `function`: the input block passes through (:1347-1348)
    n_calls = len(ms.calls)
    sess.audio_callback(indata[:, :1], outdata)
    assert len(ms.calls) == n_calls and np.allclose(outdata, 0.25)
