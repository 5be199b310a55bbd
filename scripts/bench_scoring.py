"""Scoring benchmark (BASELINE config 3): encode a synthetic 100k-news corpus once, then score impressions.

Prints one JSON line with news-enc/s (CNE forward, eval) and scored impressions/s (SUE + dot product from the
cached corpus, 50 history + 37 candidates per impression)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--news', type=int, default=100000)
    ap.add_argument('--impressions', type=int, default=4096)
    ap.add_argument('--candidates', type=int, default=37)
    ap.add_argument('--chunk', type=int, default=4096)
    ap.add_argument('--batch', type=int, default=256)
    a = ap.parse_args()
    import nnr_b200
    from nnr_b200.scoring import CorpusScorer
    from nnr_b200.synthetic import SyntheticMIND
    from bench import make_config
    class A: pass
    args = A(); args.vocab = 40000; args.dropout = 0.2
    cfg = make_config(args)
    syn = SyntheticMIND(news_num=a.news, vocabulary_size=40000, lengths='mind', seed=0)
    cfg.pretrained_word_embedding = syn.word_table()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    m = nnr_b200.Model(cfg); m.initialize(); m.to(dev).eval()
    sc = CorpusScorer(m, syn.news_title_text, syn.news_title_mask, syn.news_abstract_text, syn.news_abstract_mask,
                      syn.news_category, syn.news_subCategory, chunk=a.chunk)
    sc.encode_corpus(); torch.cuda.synchronize()                      # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sc.encode_corpus(); e1.record(); torch.cuda.synchronize()
    enc_ms = e0.elapsed_time(e1)
    hist, hl, cand = syn.sample_behaviors(a.impressions, news_num=a.candidates, seed=1)
    hist, hl, cand = torch.from_numpy(hist).to(dev), torch.from_numpy(hl).to(dev), torch.from_numpy(cand).to(dev)
    def run():
        out = []
        for i in range(0, a.impressions, a.batch):
            out.append(sc.score(hist[i:i + a.batch], hl[i:i + a.batch], cand[i:i + a.batch]))
        return out
    run(); torch.cuda.synchronize()
    e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
    sc_ms = e0.elapsed_time(e1)
    tokens = int(syn.title_len.sum() + syn.abstract_len.sum())
    print(json.dumps({'metric': 'cne_scoring_news_enc_per_sec', 'value': a.news / (enc_ms / 1e3), 'unit': 'news/s',
                      'encode_ms': enc_ms, 'corpus_news': a.news, 'corpus_tokens': tokens, 'chunk': a.chunk,
                      'scored_impressions_per_sec': a.impressions / (sc_ms / 1e3), 'candidates_per_impression': a.candidates,
                      'score_ms': sc_ms, 'finite': bool(torch.isfinite(torch.cat([o.reshape(-1) for o in out])).all())}))


if __name__ == '__main__':
    main()
