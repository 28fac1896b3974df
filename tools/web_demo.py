"""Minimal chat web UI (single HTML page + the OpenAI-compatible endpoint of ``tools/openai_api.py``); replaces the
reference's streamlit demos (``web_demo.py`` / ``web_demo_internlm.py``) without extra dependencies.

    python tools/web_demo.py --ckpt_dir llm_ckpts/1000 --tokenizer tokenizer.model --port 8080
"""
from fastapi.responses import HTMLResponse
from openai_api import app, main

PAGE = """<!doctype html><meta charset=utf-8><title>internevo_b200 chat</title>
<style>body{font-family:sans-serif;max-width:760px;margin:2em auto}#log div{margin:.5em 0;white-space:pre-wrap}
.u{color:#035}.b{color:#252}</style><h3>internevo_b200 chat</h3><div id=log></div>
<form onsubmit="send();return false"><input id=q style="width:85%" autofocus><button>send</button></form>
<script>
const msgs=[];async function send(){const q=document.getElementById('q');const t=q.value;q.value='';if(!t)return;
msgs.push({role:'user',content:t});add('u','User: '+t);const el=add('b','Bot: ');
const r=await fetch('/v1/chat/completions',{method:'POST',headers:{'content-type':'application/json'},
body:JSON.stringify({messages:msgs,stream:true})});const rd=r.body.getReader();const dec=new TextDecoder();let acc='';
for(;;){const {done,value}=await rd.read();if(done)break;for(const line of dec.decode(value).split('\\n')){
if(!line.startsWith('data: ')||line.includes('[DONE]'))continue;acc+=JSON.parse(line.slice(6)).choices[0].delta.content||'';
el.textContent='Bot: '+acc;}}msgs.push({role:'assistant',content:acc});}
function add(c,t){const d=document.createElement('div');d.className=c;d.textContent=t;document.getElementById('log').appendChild(d);return d}
</script>"""


@app.get("/", response_class=HTMLResponse)
def index():
    return PAGE


if __name__ == "__main__":
    main()
