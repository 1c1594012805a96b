"""GPU: tokenise -> pre-encoded shard -> train from the shard (scripts/pre_encode.py:425-511 writer,
training/data.py:561-573 + train_muse.py:689-690 consumer)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from open_muse_b200 import MaskGitTransformer, MaskGitVQGAN  # noqa: E402
from open_muse_b200.pre_encode import PreEncodedShardWriter, collate_pre_encoded, iter_pre_encoded, pre_encode_images  # noqa: E402

DEV = "cuda"
VAE, CLIP = "openMUSE/maskgit-vqgan-imagenet-f16-256", "openMUSE/clip"


def test_tokenise_write_read_train(tmp_path):
    torch.manual_seed(0)
    vq = MaskGitVQGAN(resolution=32, hidden_channels=32, channel_mult=(1, 2), num_res_blocks=1, z_channels=16,
                      num_embeddings=64, quantized_embed_dim=16).to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    imgs = [torch.rand(6, 3, 32, 32, generator=g) for _ in range(3)]
    ehs = [torch.randn(6, 5, 32, generator=g) for _ in range(3)]
    batches = [([f"s{b}_{i}" for i in range(6)], imgs[b], ehs[b], [{"i": i} for i in range(6)]) for b in range(3)]
    path = str(tmp_path / "00000.tar")
    with PreEncodedShardWriter(path, VAE, CLIP) as w:
        assert pre_encode_images(vq, batches, w) == 18
    samples = list(iter_pre_encoded(path, VAE, CLIP))
    batch = collate_pre_encoded(samples)
    direct = torch.cat([vq.get_code(x.to(DEV)) for x in imgs]).cpu()
    assert torch.equal(batch["image_input_ids"], direct)            # ids survive the wire format bit for bit
    assert torch.equal(batch["encoder_hidden_states"], torch.cat(ehs))
    # train a text-conditional transformer straight from the shard (the is_pre_encode branch of train_muse.py)
    m = MaskGitTransformer(vocab_size=65, hidden_size=64, num_hidden_layers=2, num_attention_heads=1, intermediate_size=128,
                           max_position_embeddings=256, codebook_size=64, num_vq_tokens=256, add_cross_attention=True,
                           encoder_hidden_size=32, hidden_dropout=0.0, attention_dropout=0.0).to(DEV).train()
    opt = torch.optim.AdamW(m.parameters(), lr=2e-3)
    ids = batch["image_input_ids"].to(DEV)
    enc = batch["encoder_hidden_states"].to(DEV)
    mask = torch.rand(ids.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2)) < 0.5
    inp, lab = torch.where(mask, 64, ids), torch.where(mask, ids, -100)
    losses = []
    for _ in range(6):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, loss = m(inp, encoder_hidden_states=enc, labels=lab)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(float(loss))
    assert losses[-1] < losses[0]
