"""Checkpoint compatibility of the BERT model with stock Hugging-Face ``BertForSequenceClassification`` (CPU):
HF state_dict -> ours (packed QKV) -> same logits, with and without an attention mask; ours -> HF loads strictly."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _pair():
    from baton_b200.models.bert import BertConfig, BertForSequenceClassification
    torch.manual_seed(0)
    hf_cfg = transformers.BertConfig(vocab_size=211, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                     intermediate_size=128, max_position_embeddings=32, type_vocab_size=2,
                                     hidden_act="gelu_pytorch_tanh", hidden_dropout_prob=0.0,
                                     attention_probs_dropout_prob=0.0, num_labels=3)
    hf = transformers.BertForSequenceClassification(hf_cfg).eval()
    ours = BertForSequenceClassification(BertConfig(vocab_size=211, hidden_size=64, num_hidden_layers=2,
                                                    num_attention_heads=4, intermediate_size=128,
                                                    max_position_embeddings=32, num_labels=3))
    nn_train = torch.nn.Module.train
    nn_train(ours, False)
    return hf, ours


def test_hf_state_dict_loads_and_logits_match():
    hf, ours = _pair()
    ours.load_hf_state_dict(hf.state_dict())
    ids = torch.randint(0, 211, (3, 16))
    mask = torch.ones(3, 16, dtype=torch.long)
    mask[1, 9:] = 0
    mask[2, 4:] = 0
    types = torch.zeros(3, 16, dtype=torch.long)
    types[:, 8:] = 1
    with torch.no_grad():
        ref = hf(input_ids=ids).logits
        got = ours(ids)
        assert torch.allclose(got, ref, atol=2e-4, rtol=1e-3), (got, ref)
        ref_m = hf(input_ids=ids, attention_mask=mask, token_type_ids=types).logits
        got_m = ours(ids, attention_mask=mask, token_type_ids=types)
        assert torch.allclose(got_m, ref_m, atol=2e-4, rtol=1e-3), (got_m, ref_m)
        assert float((got_m - got).abs().max()) > 1e-5          # the mask / token types actually change the result


def test_our_checkpoint_loads_into_stock_hf_model():
    hf, ours = _pair()
    missing = hf.load_state_dict(ours.hf_state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ids = torch.randint(0, 211, (2, 12))
    with torch.no_grad():
        assert torch.allclose(ours(ids), hf(input_ids=ids).logits, atol=2e-4, rtol=1e-3)
    # round trip is exact
    again = type(ours)(ours.config)
    again.load_hf_state_dict(ours.hf_state_dict())
    for (k, a), (_, b) in zip(ours.state_dict().items(), again.state_dict().items()):
        assert torch.equal(a, b), k
