"""CPU: the pre-encoded shard wire format (scripts/pre_encode.py:136-243 writer, training/data.py:561-573 reader) --
tar member naming, torch.save payloads, webdataset's key / extension rules -- round-trips through the package's writer
and reader and is readable with nothing but ``tarfile`` + ``torch.load`` (what webdataset's torch_loads does)."""
import io
import json
import tarfile

import pytest
import torch

from open_muse_b200.pre_encode import PreEncodedShardWriter, checkpoint_ext, collate_pre_encoded, iter_pre_encoded

VAE, CLIP = "openMUSE/vqgan-f16-8192-laion", "openMUSE/CLIP-ViT-L-14-DataComp.XL-s13B-b90K-penultimate"


def test_extension_names_match_the_reference_constants():
    assert checkpoint_ext(VAE) == "openMUSE.vqgan-f16-8192-laion.pth"                      # scripts/pre_encode.py:54-56
    assert checkpoint_ext(CLIP) == "openMUSE.CLIP-ViT-L-14-DataComp.XL-s13B-b90K-penultimate.pth"


def test_shard_round_trip(tmp_path):
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 8192, (5, 256), generator=g)
    ehs = torch.randn(5, 77, 768, generator=g)
    path = str(tmp_path / "00000.tar")
    with PreEncodedShardWriter(path, VAE, CLIP, mtime=0) as w:
        for i in range(5):
            w.write(f"sample{i:04d}", ids[i], ehs[i], {"caption": f"c{i}", "attention_mask_length": 7 + i})
        with pytest.raises(ValueError):
            w.write("bad.key", ids[0], ehs[0])
        with pytest.raises(TypeError):
            w.write("k", ids[0].int(), ehs[0])
    # raw layout: three members per sample, contiguous, webdataset's TarWriter attributes
    with tarfile.open(path) as t:
        names = [m.name for m in t.getmembers()]
        assert names[:3] == ["sample0000." + checkpoint_ext(VAE), "sample0000." + checkpoint_ext(CLIP), "sample0000.json"]
        assert len(names) == 15 and all(m.mode == 0o444 and m.uname == "bigdata" for m in t.getmembers())
        raw = torch.load(io.BytesIO(t.extractfile(names[0]).read()))  # what wds.autodecode.torch_loads does
        assert raw.dtype == torch.int64 and torch.equal(raw, ids[0])
        assert json.loads(t.extractfile("sample0003.json").read())["attention_mask_length"] == 10
    got = list(iter_pre_encoded(path, VAE, CLIP, keep_metadata=True))
    assert [s["__key__"] for s in got] == [f"sample{i:04d}" for i in range(5)]
    batch = collate_pre_encoded(got)
    assert torch.equal(batch["image_input_ids"], ids) and torch.equal(batch["encoder_hidden_states"], ehs)
    assert got[2]["json"]["caption"] == "c2"
    # class-conditional / ids-only shards and streamed file objects
    buf = io.BytesIO()
    with PreEncodedShardWriter(buf, VAE) as w:
        w.write("a", ids[0])
        w.write("b", ids[1], metadata={"x": 1})
    buf.seek(0)
    only = list(iter_pre_encoded(buf, VAE))
    assert len(only) == 2 and torch.equal(only[1]["image_input_ids"], ids[1]) and "json" not in only[1]
    # a reader configured for other checkpoints finds nothing usable (incomplete samples are skipped)
    assert list(iter_pre_encoded(path, "someone/else", CLIP)) == []
