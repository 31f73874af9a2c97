"""One-box multi-GPU dispatch behind the `RapidOcrModel` surface (north_star: "pages and cropped text-lines ...
sharded round-robin across the GPUs of one 8xB200 box").

`B200OcrPool` owns one `B200OcrModel` (det + rec engines) per device and one worker thread per device; the C-ABI calls
release the GIL, so the GPUs run concurrently from one Python process — the shape RapidDoc's single-process orchestration
(`AtomModelSingleton`, rapid_doc/backend/pipeline/model_init.py:57-88) needs.  There is no data-path collective: pages and
text-line batches are independent units; results come back in input order.

  det_batch_predict(pages)      pages dealt round-robin (parallel.round_robin)
  ocr(crops, det=False)         the reference's global plan (sort by w/h, consecutive batches of rec_batch_num,
                                rapid_ocr.py:411-440) is made ONCE, whole batches are dealt round-robin, so every crop is
                                recognised in the same batch, padded to the same width, as on one GPU
  ocr_pages(pages)              page-parallel det+rec; each device batches the crops of ITS pages (as the reference batches
                                the crops of whatever window one recogniser is handed)

`model_factory(device)` is injectable so the dispatch logic is testable on CPU with stand-in models.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .parallel import round_robin


class B200OcrPool:
    def __init__(self, devices, model_factory=None, **model_kwargs):
        devices = list(devices)
        assert devices, "at least one device"
        if model_factory is None:
            from .ocr import B200OcrModel

            def model_factory(dev):
                return B200OcrModel(device=dev, **model_kwargs)
        self.devices = devices
        self.models = [model_factory(d) for d in devices]
        self._workers = ThreadPoolExecutor(max_workers=len(devices))
        m0 = self.models[0]
        # the attributes the reference's callers read (rapid_ocr.py:43-162)
        self.drop_score = getattr(m0, "drop_score", 0.5)
        self.enable_merge_det_boxes = getattr(m0, "enable_merge_det_boxes", True)
        self.is_seal = False
        self.text_detector = getattr(m0, "text_detector", None)
        self.text_recognizer = getattr(m0, "text_recognizer", None)
        self.rec_batch_num = getattr(m0, "rec_batch_num", 6)

    def _scatter(self, shares, fn):
        """shares[r] = list of unit indices for device r; fn(model, indices) -> list of results; gathered in input order."""
        futs = [self._workers.submit(fn, m, idx) if len(idx) else None for m, idx in zip(self.models, shares)]
        total = sum(len(s) for s in shares)
        out = [None] * total
        for idx, f in zip(shares, futs):
            if f is None:
                continue
            for i, r in zip(idx, f.result()):
                out[i] = r
        return out

    # ---- rapid_ocr.py:474-497
    def det_batch_predict(self, img_list, max_batch_size=8):
        if img_list is None or len(img_list) == 0:
            return []
        world = len(self.models)
        shares = [round_robin(len(img_list), world, r) for r in range(world)]
        return self._scatter(shares, lambda m, idx: m.det_batch_predict([img_list[i] for i in idx], max_batch_size))

    def ocr_pages(self, pages, mfd_res_list=None, drop_score=None):
        world = len(self.models)
        shares = [round_robin(len(pages), world, r) for r in range(world)]
        return self._scatter(shares, lambda m, idx: m.ocr_pages([pages[i] for i in idx],
                                                               [mfd_res_list[i] for i in idx] if mfd_res_list else None, drop_score))

    def rec_shares(self, shapes):
        """Whole reference batches dealt round-robin: [indices for device r], each in ascending w/h order."""
        world = len(self.models)
        ratios = np.array([w / float(h) for h, w in shapes])
        order = np.argsort(ratios)
        bs = int(self.rec_batch_num)
        shares = [[] for _ in range(world)]
        for k, b0 in enumerate(range(0, len(order), bs)):
            shares[k % world].extend(int(i) for i in order[b0: b0 + bs])
        return shares

    # ---- rapid_ocr.py:225-299
    def ocr(self, img, det=True, rec=True, **kw):
        if isinstance(img, list) and not det and rec and len(img) > self.rec_batch_num and not kw.get("return_word_box"):
            shares = self.rec_shares([c.shape[:2] for c in img])
            res = self._scatter(shares, lambda m, idx: m.ocr([img[i] for i in idx], det=False, rec=True)[0])
            return [res]
        return self.models[0].ocr(img, det=det, rec=rec, **kw)

    def __call__(self, img, mfd_res=None):
        return self.models[0](img, mfd_res=mfd_res)

    def text_recognizer_call(self, args, **kw):
        return self.models[0].text_recognizer_call(args, **kw)

    def close(self):
        self._workers.shutdown(wait=True)
