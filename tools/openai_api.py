"""OpenAI-compatible chat-completions server over a training checkpoint (reference ``tools/openai_api.py``).

    python tools/openai_api.py --ckpt_dir llm_ckpts/1000 --tokenizer tokenizer.model --port 8000
    curl localhost:8000/v1/chat/completions -d '{"model":"internlm2","messages":[{"role":"user","content":"hi"}]}'
"""
import argparse
import json
import time
import uuid
from typing import List, Optional

from fastapi import FastAPI
from fastapi.responses import StreamingResponse
from pydantic import BaseModel

app = FastAPI()
STATE = {}


class Message(BaseModel):
    role: str
    content: str


class ChatRequest(BaseModel):
    model: str = "internlm2"
    messages: List[Message]
    temperature: float = 0.8
    top_p: float = 0.8
    max_tokens: Optional[int] = 512
    stream: bool = False


def build_prompt(messages: List[Message]) -> str:
    """InternLM chat template: <|User|>:…<eoh>\\n<|Bot|>:…<eoa>\\n"""
    out = []
    for m in messages:
        if m.role == "system":
            out.append(f"<|System|>:{m.content}\n")
        elif m.role == "user":
            out.append(f"<|User|>:{m.content}<eoh>\n")
        else:
            out.append(f"<|Bot|>:{m.content}<eoa>\n")
    out.append("<|Bot|>:")
    return "".join(out)


def _stream(prompt, req):
    from load_internlm_model import internlm_interactive_generation

    return internlm_interactive_generation(STATE["model"], STATE["tokenizer"], prompt, max_new_tokens=req.max_tokens or 512,
                                           temperature=max(req.temperature, 1e-3), top_p=req.top_p,
                                           do_sample=req.temperature > 0,
                                           additional_eos_token_list=STATE.get("extra_eos"))


@app.get("/v1/models")
def models():
    return {"object": "list", "data": [{"id": STATE.get("name", "internlm2"), "object": "model"}]}


@app.post("/v1/chat/completions")
def chat(req: ChatRequest):
    prompt = build_prompt(req.messages)
    cid, created = f"chatcmpl-{uuid.uuid4().hex}", int(time.time())
    if req.stream:
        def sse():
            prev = ""
            for text in _stream(prompt, req):
                delta, prev = text[len(prev):], text
                chunk = {"id": cid, "object": "chat.completion.chunk", "created": created, "model": req.model,
                         "choices": [{"index": 0, "delta": {"content": delta}, "finish_reason": None}]}
                yield f"data: {json.dumps(chunk, ensure_ascii=False)}\n\n"
            yield "data: [DONE]\n\n"

        return StreamingResponse(sse(), media_type="text/event-stream")
    text = ""
    for text in _stream(prompt, req):
        pass
    return {"id": cid, "object": "chat.completion", "created": created, "model": req.model,
            "choices": [{"index": 0, "message": {"role": "assistant", "content": text}, "finish_reason": "stop"}]}


def main():
    import sentencepiece as spm
    import uvicorn
    from load_internlm_model import initialize_internlm_model

    p = argparse.ArgumentParser()
    p.add_argument("--model_type", default="INTERNLM2_PUBLIC")
    p.add_argument("--ckpt_dir", required=True)
    p.add_argument("--tokenizer", required=True)
    p.add_argument("--host", default="0.0.0.0")
    p.add_argument("--port", type=int, default=8000)
    a = p.parse_args()
    sp = spm.SentencePieceProcessor()
    sp.Load(a.tokenizer)
    STATE.update(model=initialize_internlm_model(a.model_type, a.ckpt_dir), tokenizer=sp, name=a.model_type.lower())
    uvicorn.run(app, host=a.host, port=a.port)


if __name__ == "__main__":
    main()
