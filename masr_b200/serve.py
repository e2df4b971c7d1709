"""The message protocol of the reference's streaming websocket endpoint (infer_server.py:103-156) for MANY concurrent
connections on one GPU.

The reference serves one stream at a time from one global predictor: every binary message is raw 16 kHz mono int16 PCM; a
message ending in ``b'end'`` closes the utterance; after every non-empty message the server answers
``{"code": 0, "result": <text so far>}`` (the text only changes when ``predict_stream`` returned a result), and
``{"code": 1, "msg": "recognition fail, no resource!"}`` when no predictor is free.  ``StreamSessions`` keeps that contract
per connection over a ``StreamPool``: the messages that arrived for different connections in one tick are decoded together
(one batched chunk step per round).  Transport (FastAPI / websockets) stays outside: this class maps messages to replies.
"""
from __future__ import annotations

from typing import Dict, List, Optional

END_MARK = b"end"
NO_RESOURCE = {"code": 1, "msg": "recognition fail, no resource!"}
FAILED = {"code": 2, "msg": "recognition fail!"}


class StreamSessions:
    def __init__(self, pool):
        """``pool``: a ``masr_b200.stream_pool.StreamPool`` (its slots are the "predictors" of the reference server)."""
        self.pool = pool
        self.free: List[int] = list(range(pool.S))[::-1]
        self.text: Dict[int, str] = {}

    def open(self) -> Optional[int]:
        """A new connection -> session id, or None when every slot is busy (answer ``NO_RESOURCE`` and close)."""
        if not self.free:
            return None
        s = self.free.pop()
        self.pool.reset_stream(s)
        self.text[s] = ""
        return s

    def close(self, session: int):
        """Connection closed (by the client or after ``end``): reset the stream and free the slot (infer_server.py:139-141)."""
        if session in self.text:
            del self.text[session]
            self.pool.reset_stream(session)
            self.free.append(session)

    def feed(self, messages: Dict[int, bytes]) -> Dict[int, dict]:
        """One binary message per session -> one reply per session that sent a non-empty message.  A session whose message
        ended in ``b'end'`` is closed after its reply, as the reference closes the websocket."""
        mid, last = {}, {}
        for s, data in messages.items():
            if s not in self.text:
                raise KeyError(f"unknown session {s}")
            if len(data) == 0:
                continue                                   # infer_server.py:113: ignored, no reply
            if data[-3:] == END_MARK:
                last[s] = data[:-3]
            else:
                mid[s] = data
        replies: Dict[int, dict] = {}
        for grp, is_end in ((mid, False), (last, True)):
            if not grp:
                continue
            try:
                out = self.pool.push(grp, is_end=is_end, on_error="return")
                bad = set(getattr(self.pool, "last_errors", {}))
            except Exception:                               # infer_server.py:130-137 (a failure of the whole batched step)
                for s in grp:
                    replies[s] = dict(FAILED)
                continue
            for s in grp:
                if s in bad:                                # only the offending connection fails (its stream state is untouched)
                    replies[s] = dict(FAILED)
                    continue
                if out.get(s) is not None:
                    self.text[s] = out[s]["text"]
                replies[s] = {"code": 0, "result": self.text[s]}
        for s in last:
            self.close(s)
        return replies
