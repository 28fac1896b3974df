"""Checkpoint storage: ``local:/path`` plus pluggable object stores (``boto3:s3://…``, ``volc:vc://…``, ``oss2:…``) with
optional asynchronous upload (reference ``internlm/utils/storage_manager.py:95-1288``).

Public surface kept: ``llm_save / llm_load / get_fns / check_folder``, ``init_storage_manager``,
``get_storage_manager().wait()``, ``try_get_storage_backend``.  Object-store SDKs are imported lazily (none is present in
this image); each backend (``LocalClient``, ``Boto3Client``, ``VolcClient``, ``AliClient``) only has to provide ``upload / download / list / exists / delete`` on raw bytes, the manager
owns serialisation (``torch.save``), the ``/dev/shm`` staging folder, md5 sidecars, the thread pool and the ``{step}.step``
completion marker that is written only after every asynchronous upload of the step has finished.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import io
import os
import re
import shutil
import socket
import stat
from typing import Any, Callable, Dict, List, Optional

import torch

from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)
_PREFIX = re.compile(r"^(boto3|volc|oss2|local):")


def try_get_storage_backend(path: str):
    """``"boto3:s3://bucket/x"`` → ``("boto3", "s3://bucket/x")``; bare paths are local."""
    m = _PREFIX.match(path)
    if m:
        return m.group(1), path[m.end():]
    for scheme, backend in (("s3://", "boto3"), ("vc://", "volc"), ("ali://", "oss2")):   # URL scheme names the backend
        if path.startswith(scheme):
            return backend, path
    if os.environ.get("RANK", "0") == "0":
        logger.warning(f"path: '{path}' not start with backend prefix, guess it is the backend of local.")
    return "local", path


def compute_file_md5_by_chunk(file_name: str) -> str:
    h = hashlib.md5()
    with open(file_name, "rb") as f:
        for chunk in iter(lambda: f.read(4 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


class StorageClient:
    """Backend interface."""

    def upload_file(self, local_path: str, remote: str) -> None:
        raise NotImplementedError

    def upload_bytes(self, data: bytes, remote: str) -> None:
        raise NotImplementedError

    def download_bytes(self, remote: str) -> bytes:
        raise NotImplementedError

    def list(self, remote: str) -> List[str]:
        raise NotImplementedError

    def exists(self, remote: str) -> bool:
        raise NotImplementedError

    def delete(self, remote: str) -> None:
        raise NotImplementedError

    # ---- object-level verbs (what the reference's clients expose one by one per backend, ``storage_manager.py:335-845``),
    # written ONCE on top of the byte primitives above; ``fp`` is the path in the form this client takes
    def sync_upload_fileobj(self, fp: str, saved_obj=None, **kwargs) -> None:
        assert saved_obj is not None, "saved_obj is None!"
        buf = io.BytesIO()
        torch.save(saved_obj, buf, **kwargs)
        self.upload_bytes(buf.getvalue(), fp)

    def async_upload_fileobj(self, fp: str, local_nvme_path: str) -> str:
        """Second half of an asynchronous save: the object already sits in the staging file; write its md5 sidecar, upload it,
        drop the staging file.  Runs on a worker thread of the manager's pool."""
        self.upload_bytes(compute_file_md5_by_chunk(local_nvme_path).encode(), fp + ".md5")
        self.upload_file(local_nvme_path, fp)
        if os.path.exists(local_nvme_path):
            os.remove(local_nvme_path)
        return fp

    def load(self, fp: str, **kwargs):
        kwargs.setdefault("map_location", "cpu")
        kwargs.setdefault("weights_only", False)
        return torch.load(io.BytesIO(self.download_bytes(fp)), **kwargs)

    def is_fp_exists(self, fp: str) -> bool:
        return self.exists(fp) or len(self.list(fp)) > 0

    def assert_fp_exists(self, fp: str) -> None:
        assert self.is_fp_exists(fp), f"'{fp}' does not exist"

    def get_fns(self, fp: str) -> List[str]:
        return self.list(fp)

    def delete_obj(self, fp: str) -> None:
        self.delete(fp)


class LocalClient(StorageClient):
    def upload_file(self, local_path, remote):
        os.makedirs(os.path.dirname(remote) or ".", exist_ok=True)
        shutil.move(local_path, remote)

    def upload_bytes(self, data, remote):
        os.makedirs(os.path.dirname(remote) or ".", exist_ok=True)
        tmp = remote + ".tmp"
        with open(tmp, "wb") as f:
            f.write(data)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, remote)

    def download_bytes(self, remote):
        with open(remote, "rb") as f:
            return f.read()

    def list(self, remote):
        if not os.path.exists(remote):
            return []
        if os.path.isfile(remote):
            return [os.path.basename(remote)]
        return sorted(os.listdir(remote))

    def exists(self, remote):
        return os.path.exists(remote)

    def delete(self, remote):
        if os.path.isdir(remote):
            shutil.rmtree(remote, ignore_errors=True)
        elif os.path.exists(remote):
            os.remove(remote)


class Boto3Client(StorageClient):
    """``boto3:s3://{bucket}.{endpoint}/{key}`` — credentials from ``S3_ACCESS_KEY_ID`` / ``S3_SECRET_ACCESS_KEY_ID``."""

    def __init__(self, endpoint: str):
        import boto3  # noqa: lazy
        import botocore

        self.client = boto3.client(
            "s3", endpoint_url=endpoint, aws_access_key_id=os.environ["S3_ACCESS_KEY_ID"],
            aws_secret_access_key=os.environ["S3_SECRET_ACCESS_KEY_ID"], use_ssl=False,
            config=botocore.config.Config(retries={"max_attempts": 5}))

    @staticmethod
    def split(remote: str):
        m = re.match(r"^s3://([^/.]+)\.?([^/]*)/(.*)$", remote)
        assert m, f"bad s3 path {remote}"
        return m.group(1), m.group(3)

    def upload_file(self, local_path, remote):
        b, k = self.split(remote)
        self.client.upload_file(local_path, b, k)

    def upload_bytes(self, data, remote):
        b, k = self.split(remote)
        self.client.upload_fileobj(io.BytesIO(data), b, k)

    def download_bytes(self, remote):
        b, k = self.split(remote)
        buf = io.BytesIO()
        self.client.download_fileobj(b, k, buf)
        return buf.getvalue()

    def list(self, remote):
        b, k = self.split(remote)
        k = k.rstrip("/") + "/"
        names = set()
        for page in self.client.get_paginator("list_objects_v2").paginate(Bucket=b, Prefix=k):
            for o in page.get("Contents", []):
                names.add(o["Key"][len(k):].split("/")[0])
        return sorted(names)

    def exists(self, remote):
        b, k = self.split(remote)
        return self.client.list_objects_v2(Bucket=b, Prefix=k, MaxKeys=1).get("KeyCount", 0) > 0

    def delete(self, remote):
        b, k = self.split(remote)
        self.client.delete_object(Bucket=b, Key=k)


def _credentials(prefix: str):
    """``ACCESS_KEY`` / ``SECRET_ACCESS_KEY`` win over ``{prefix}_ACCESS_KEY_ID`` / ``{prefix}_SECRET_ACCESS_KEY_ID``."""
    ak = os.environ.get("ACCESS_KEY") or os.environ.get(f"{prefix}_ACCESS_KEY_ID")
    sk = os.environ.get("SECRET_ACCESS_KEY") or os.environ.get(f"{prefix}_SECRET_ACCESS_KEY_ID")
    assert ak and sk, f"set {prefix}_ACCESS_KEY_ID / {prefix}_SECRET_ACCESS_KEY_ID (or ACCESS_KEY / SECRET_ACCESS_KEY)"
    return ak, sk


def _split_bucket_url(remote: str, scheme: str):
    """``{scheme}{bucket}.{endpoint}/{key}`` → ``(bucket, endpoint, key)``."""
    m = re.match(rf"^{re.escape(scheme)}([^/.]+)\.([^/]+)/?(.*)$", remote)
    assert m, f"url '{remote}' is not a valid {scheme} url: expected {scheme}<bucket>.<endpoint>/<key>"
    return m.group(1), m.group(2), m.group(3)


def _under(key: str, prefix: str) -> bool:
    """Object stores match raw string prefixes (``run/7`` also matches ``run/7.step``): keep the object itself and what
    lies below it as a folder."""
    prefix = prefix.rstrip("/")
    return prefix == "" or key == prefix or key.startswith(prefix + "/")


def _first_segments(keys, prefix: str) -> List[str]:
    """Names directly under ``prefix`` (files or "folders"), the listing the checkpoint code expects."""
    prefix, names = prefix.rstrip("/"), set()
    for k in keys:
        if not _under(k, prefix):
            continue
        rest = k[len(prefix):].strip("/")
        names.add(rest.split("/", 1)[0] if rest else os.path.basename(k))
    return sorted(names)


class VolcClient(StorageClient):
    """``volc:vc://{bucket}.{endpoint}/{key}`` through the ``tos`` SDK (ByteDance TOS; reference
    ``storage_manager.py:485-672``).  The region is derived from the endpoint (``tos-cn-beijing.volces.com`` →
    ``cn-beijing``)."""

    PART = 64 << 20

    def __init__(self, endpoint: str, region: Optional[str] = None):
        import tos  # noqa: lazy — not part of the image

        ak, sk = _credentials("VOLC")
        region = region or "-".join(endpoint.split(".")[0].split("-")[1:])
        self.client = tos.TosClientV2(ak, sk, endpoint, region, enable_crc=False)

    @staticmethod
    def split(remote: str):
        b, _, k = _split_bucket_url(remote, "vc://")
        return b, k

    def _keys(self, bucket: str, prefix: str):
        token = None
        while True:
            kw = dict(prefix=prefix) if token is None else dict(prefix=prefix, continuation_token=token)
            res = self.client.list_objects_type2(bucket, **kw)
            for item in getattr(res, "contents", None) or []:
                yield item.key
            if not getattr(res, "is_truncated", False):
                return
            token = res.next_continuation_token

    def upload_bytes(self, data, remote):
        b, k = self.split(remote)
        self.client.put_object(b, k, content=io.BytesIO(data))

    def upload_file(self, local_path, remote):
        b, k = self.split(remote)
        if os.path.getsize(local_path) <= self.PART:
            with open(local_path, "rb") as f:
                self.client.put_object(b, k, content=f)
            return
        upload_id, parts, n = self.client.create_multipart_upload(b, k).upload_id, [], 1
        with open(local_path, "rb") as f:
            for chunk in iter(lambda: f.read(self.PART), b""):
                parts.append(self.client.upload_part(b, k, upload_id, n, content=io.BytesIO(chunk)))
                n += 1
        self.client.complete_multipart_upload(b, k, upload_id, parts)

    def download_bytes(self, remote):
        b, k = self.split(remote)
        return self.client.get_object(b, k).read()

    def list(self, remote):
        b, k = self.split(remote)
        return _first_segments(self._keys(b, k), k)

    def exists(self, remote):
        b, k = self.split(remote)
        return any(_under(key, k) for key in self._keys(b, k))

    def delete(self, remote):
        b, k = self.split(remote)
        self.client.delete_object(b, k)


class AliClient(StorageClient):
    """``oss2:ali://{bucket}.{endpoint}/{key}`` through the ``oss2`` SDK (Aliyun OSS; reference
    ``storage_manager.py:675-816``); one ``Bucket`` handle per bucket name."""

    def __init__(self, endpoint: str):
        import oss2  # noqa: lazy — not part of the image

        self._sdk, self.endpoint = oss2, endpoint
        self.auth = oss2.Auth(*_credentials("ALI"))
        self._buckets: Dict[str, Any] = {}

    def _bucket(self, remote: str):
        b, _, k = _split_bucket_url(remote, "ali://")
        if b not in self._buckets:
            self._buckets[b] = self._sdk.Bucket(self.auth, self.endpoint, b, enable_crc=False)
        return self._buckets[b], k

    def upload_bytes(self, data, remote):
        bucket, k = self._bucket(remote)
        bucket.put_object(k, data)

    def upload_file(self, local_path, remote):
        bucket, k = self._bucket(remote)
        self._sdk.resumable_upload(bucket, k, local_path) if hasattr(self._sdk, "resumable_upload") else \
            bucket.put_object_from_file(k, local_path)

    def download_bytes(self, remote):
        bucket, k = self._bucket(remote)
        return bucket.get_object(k).read()

    def list(self, remote):
        bucket, k = self._bucket(remote)
        return _first_segments((o.key for o in self._sdk.ObjectIteratorV2(bucket, prefix=k)), k)

    def exists(self, remote):
        bucket, k = self._bucket(remote)
        return any(_under(o.key, k) for o in self._sdk.ObjectIteratorV2(bucket, prefix=k))

    def delete(self, remote):
        bucket, k = self._bucket(remote)
        bucket.delete_object(k)


def get_tmp_file_name(tmp_local_folder: str, fp: str) -> str:
    """Staging file of an asynchronous upload: unique per host, process and remote path (reference ``:840-856``)."""
    base = re.sub(r"^[a-z0-9]+://", "", fp).replace(os.path.sep, "_")
    return os.path.join(tmp_local_folder, f"{socket.gethostname()}-{os.getpid()}-{base}")


def get_mount_point_free_size(path: str) -> float:
    """Free space of the file system holding ``path`` in GB."""
    st = os.statvfs(path)
    return st.f_bavail * st.f_frsize / (1 << 30)


def check_tmp_folder_accessibility(tmp_local_folder: str, min_free_gb: float = 0.1):
    """The staging folder must be readable, writable, listable and not (nearly) full (reference ``:1217-1245``)."""
    os.makedirs(tmp_local_folder, exist_ok=True)
    ok = os.access(tmp_local_folder, os.R_OK | os.W_OK | os.X_OK)
    free = get_mount_point_free_size(tmp_local_folder)
    if not ok or free < min_free_gb:
        raise RuntimeError(f"async upload staging folder {tmp_local_folder}: accessible={ok}, free={free:.2f} GB "
                           f"(need >= {min_free_gb} GB)")


# ---------------------------------------------------------------------------------------------------------------------
# Parsed storage paths ("meta info").  One record type: which backend, bucket, endpoint (region for TOS), object key and - for
# asynchronous saves - the staging file the object is serialised to first.  The four reference names (``Boto3MetaInfo``,
# ``VolcMetaInfo``, ``AliMetaInfo``, ``LocalMetaInfo``; reference ``storage_manager.py:142-301,859-934``) are the same record
# with the backend fixed, and ``get_*_meta`` are its parsers; ``StorageManager.get_meta`` hands them out with the client set.
# ---------------------------------------------------------------------------------------------------------------------
class PathMetaInfo:
    backend = "local"
    scheme = ""

    def __init__(self, is_async: bool = False, handler: Optional[StorageClient] = None, bucket_name: Optional[str] = None,
                 endpoint: Optional[str] = None, file_path: str = "", async_upload_fn: Optional[Callable] = None,
                 local_nvme_path: Optional[str] = None, region: Optional[str] = None) -> None:
        self.client = handler
        self.bucket_name, self.endpoint, self.region = bucket_name, endpoint, region
        self.file_path = file_path
        self.is_async, self.local_nvme_path, self.async_upload_fn = is_async, local_nvme_path, async_upload_fn

    @property
    def url(self) -> str:
        """The path in the form the clients of this module take."""
        if not self.scheme:
            return self.file_path
        host = re.sub(r"^https?://", "", self.endpoint or "").split(":")[0]
        return f"{self.scheme}{self.bucket_name}.{host}/{self.file_path}"

    def save_args(self):
        head = (self.client, self.bucket_name, self.file_path) if self.scheme else (self.file_path,)
        return (*head, self.local_nvme_path) if (self.is_async and self.scheme) else head

    def nosave_args(self):
        return (self.client, self.bucket_name, self.file_path) if self.scheme else (self.file_path,)

    def __str__(self) -> str:
        return (f"backend: {self.backend}, is_async: {self.is_async}, bucket_name: {self.bucket_name}, endpoint: {self.endpoint}, "
                f"file_path: {self.file_path}, local_nvme_path: {self.local_nvme_path}")


class Boto3MetaInfo(PathMetaInfo):
    backend, scheme = "boto3", "s3://"
    unpack_boto3_save_meta = staticmethod(PathMetaInfo.save_args)
    unpack_boto3_nosave_meta = staticmethod(PathMetaInfo.nosave_args)


class VolcMetaInfo(PathMetaInfo):
    backend, scheme = "volc", "vc://"
    unpack_volc_save_meta = staticmethod(PathMetaInfo.save_args)
    unpack_volc_nosave_meta = staticmethod(PathMetaInfo.nosave_args)


class AliMetaInfo(PathMetaInfo):
    backend, scheme = "oss2", "ali://"
    unpack_ali_save_meta = staticmethod(PathMetaInfo.save_args)
    unpack_ali_nosave_meta = staticmethod(PathMetaInfo.nosave_args)


class LocalMetaInfo(PathMetaInfo):
    def __init__(self, file_path: str) -> None:
        super().__init__(file_path=file_path)

    unpack_local_save_meta = staticmethod(PathMetaInfo.save_args)
    unpack_local_nosave_meta = staticmethod(PathMetaInfo.nosave_args)


def unpack_save_meta(meta: PathMetaInfo):
    return meta.save_args()


def unpack_nosave_meta(meta: PathMetaInfo):
    return meta.nosave_args()


def _object_meta(cls, fp: str, tmp_local_folder: Optional[str], is_async: bool, **extra) -> PathMetaInfo:
    assert fp.startswith(cls.scheme), f"Path '{fp}' is not a {cls.backend} url ({cls.scheme}<bucket>.<endpoint>/<key>)"
    bucket, endpoint, key = _split_bucket_url(fp, cls.scheme)
    staging = get_tmp_file_name(tmp_local_folder, fp) if is_async else None
    return cls(is_async=is_async, bucket_name=bucket, endpoint=endpoint, file_path=key, local_nvme_path=staging, **extra)


def get_boto3_meta(fp: str, tmp_local_folder: str, is_async: bool) -> Boto3MetaInfo:
    meta = _object_meta(Boto3MetaInfo, fp, tmp_local_folder, is_async)
    meta.endpoint = f"http://{meta.endpoint}:80" if ":" not in meta.endpoint else f"http://{meta.endpoint}"
    return meta


def get_volc_meta(fp: str, tmp_local_folder: str, is_async: bool) -> VolcMetaInfo:
    meta = _object_meta(VolcMetaInfo, fp, tmp_local_folder, is_async)
    meta.region = "-".join(meta.endpoint.split(".")[0].split("-")[1:])      # tos-cn-beijing.volces.com -> cn-beijing
    return meta


def get_ali_meta(fp: str, tmp_local_folder: str, is_async: bool) -> AliMetaInfo:
    return _object_meta(AliMetaInfo, fp, tmp_local_folder, is_async)


def get_local_meta(fp: str) -> LocalMetaInfo:
    assert not fp.startswith(("s3://", "vc://", "ali://")), f"Path '{fp}' is not a local path"
    return LocalMetaInfo(fp)


_META_PARSERS = {"boto3": get_boto3_meta, "volc": get_volc_meta, "oss2": get_ali_meta}


def is_rank_for_log() -> bool:
    """Storage messages are printed once per job (rank 0 of the launcher's numbering; no process group is needed)."""
    return os.environ.get("RANK", "0") == "0"


class Logger:
    """Rank-0-only facade over the module logger, the form storage code logs through (reference ``:61-88``)."""

    def info(self, mesage: str):
        if is_rank_for_log():
            logger.info(mesage)

    def warning(self, mesage: str):
        if is_rank_for_log():
            logger.warning(mesage)

    def error(self, mesage: str):
        logger.error(mesage)


def _make_client(backend: str, path: str) -> StorageClient:
    if backend == "local":
        return LocalClient()
    if backend == "boto3":
        m = re.match(r"^s3://[^/.]+\.([^/]+)/", path)
        endpoint = f"http://{m.group(1)}" if m and m.group(1) else os.environ.get("S3_ENDPOINT_URL")
        return Boto3Client(endpoint)
    if backend == "volc":
        return VolcClient(_split_bucket_url(path, "vc://")[1])
    if backend == "oss2":
        return AliClient(_split_bucket_url(path, "ali://")[1])
    raise ValueError(f"unknown storage backend {backend}")


_custom_backends: Dict[str, Callable[[str], StorageClient]] = {}


def register_backend(name: str, factory: Callable[[str], StorageClient]):
    """Plug in an object-store client (e.g. for ``volc:`` / ``oss2:`` prefixes)."""
    _custom_backends[name] = factory


class SingletonMeta(type):
    """One instance per class; a second construction with arguments is an error (reference ``storage_manager.py:967-981``)."""

    _instances = {}

    def __call__(cls, *args, **kwargs):
        if cls not in cls._instances:
            cls._instances[cls] = super().__call__(*args, **kwargs)
        else:
            assert len(args) == 0 and len(kwargs) == 0, f"{cls.__name__} is a singleton class and a instance has been created."
        return cls._instances[cls]


class StorageManager:
    """Serialises objects, routes them to the backend and tracks asynchronous uploads."""

    def __init__(self, enable_save: bool = True, tmp_local_folder: Optional[str] = None, async_mode: bool = False,
                 n_async_workers: int = 8):
        self._clients: Dict[str, StorageClient] = {}
        self.async_mode = bool(async_mode and tmp_local_folder)
        self.tmp_local_folder = tmp_local_folder
        self._futures: List[concurrent.futures.Future] = []
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=n_async_workers) if self.async_mode else None
        self._to_be_marked: Optional[str] = None
        self.latest_save_folder = None
        self.latest_save_step = 0
        self.async_task_peeding = False
        if enable_save and self.async_mode:
            os.makedirs(tmp_local_folder, exist_ok=True)
            try:
                os.chmod(tmp_local_folder, stat.S_IRWXU | stat.S_IRWXG | stat.S_IRWXO)
            except PermissionError:      # created by another user with the right mode already: checked right below
                pass
            check_tmp_folder_accessibility(tmp_local_folder)
            self.try_delete_tmpfile(tmp_local_folder)

    def _client(self, path: str):
        backend, real = try_get_storage_backend(path)
        if backend not in self._clients:
            self._clients[backend] = _custom_backends[backend](real) if backend in _custom_backends else _make_client(backend, real)
        return self._clients[backend], backend, real

    def get_meta(self, path: str, is_async: Optional[bool] = None) -> PathMetaInfo:
        """Parsed form of ``path`` with this manager's client for its backend attached."""
        backend, real = try_get_storage_backend(path)
        is_async = self.async_mode if is_async is None else is_async
        meta = get_local_meta(real) if backend == "local" else _META_PARSERS[backend](
            real, self.tmp_local_folder, is_async and backend != "local")
        meta.client = self._client(path)[0]
        return meta

    def assert_fp_exists(self, folder) -> None:
        c, _, real = self._client(folder)
        assert c.exists(real), f"{folder} does not exist"

    def get_fns(self, folder) -> List[str]:
        c, _, real = self._client(folder)
        return c.list(real)

    def is_exists(self, path) -> bool:
        c, _, real = self._client(path)
        return c.exists(real)

    def save(self, save_path: str, to_save_obj: Any, async_upload=None, **kwargs):
        c, backend, real = self._client(save_path)
        use_async = self.async_mode if async_upload is None else (async_upload and self.async_mode)
        if backend == "local" or not use_async:
            buf = io.BytesIO()
            torch.save(to_save_obj, buf, **kwargs)
            c.upload_bytes(buf.getvalue(), real)
            return
        tmp = get_tmp_file_name(self.tmp_local_folder, real)
        torch.save(to_save_obj, tmp, **kwargs)
        self.async_task_peeding = True
        self._futures.append(self._pool.submit(self._upload_and_clean, c, tmp, real))

    @staticmethod
    def _upload_and_clean(client, tmp, real):
        return client.async_upload_fileobj(real, tmp)

    def try_delete_tmpfile(self, tmp_dir: str) -> int:
        """Remove staging files left behind by processes of THIS host that no longer run (a killed job's half-written
        checkpoints fill ``/dev/shm`` otherwise); files of live processes - the other ranks share the folder - are kept.
        Returns the number of files removed (reference ``storage_manager.py:1185-1196`` deletes every ``*.tmpfile``)."""
        removed, host = 0, socket.gethostname() + "-"
        if not os.path.isdir(tmp_dir):
            return 0
        for name in os.listdir(tmp_dir):
            if not name.startswith(host):
                continue
            pid = name[len(host):].split("-", 1)[0]
            if pid.isdigit() and int(pid) != os.getpid() and os.path.exists(f"/proc/{pid}"):
                continue
            try:
                os.remove(os.path.join(tmp_dir, name))
                removed += 1
            except OSError:
                pass
        return removed

    def load(self, load_path: str, **kwargs) -> Any:
        c, _, real = self._client(load_path)
        kwargs.setdefault("map_location", "cpu")
        kwargs.setdefault("weights_only", False)
        return torch.load(io.BytesIO(c.download_bytes(real)), **kwargs)

    def delete_obj(self, fp: str):
        c, _, real = self._client(fp)
        c.delete(real)

    def async_executor(self, fn: Callable, *args, **kwargs) -> None:
        if self._pool is None:
            fn(*args, **kwargs)
        else:
            self._futures.append(self._pool.submit(fn, *args, **kwargs))

    def set_pending_marker(self, marker_path: str):
        """``{step}.step`` is written by ``wait()`` once all uploads of that step are done."""
        self._to_be_marked = marker_path

    def wait(self) -> bool:
        """Block until every outstanding upload finished; then publish the completion marker."""
        ok = True
        for f in self._futures:
            try:
                f.result()
            except Exception as e:  # pragma: no cover
                ok = False
                logger.error(f"async upload failed: {e}")
        self._futures = []
        self.async_task_peeding = False
        if ok and self._to_be_marked is not None:
            c, _, real = self._client(self._to_be_marked)
            c.upload_bytes(b"", real)
            self._to_be_marked = None
        return ok


storage_manager: Optional[StorageManager] = None


def init_storage_manager(enable_save_ckpt, async_upload_tmp_folder, async_upload, use_processpool=False):
    global storage_manager
    storage_manager = StorageManager(enable_save_ckpt, tmp_local_folder=async_upload_tmp_folder, async_mode=async_upload)
    return storage_manager


def get_storage_manager() -> StorageManager:
    global storage_manager
    if storage_manager is None:
        storage_manager = StorageManager()
    return storage_manager


def wait_async_upload_finish():
    import torch.distributed as dist

    get_storage_manager().wait()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def check_folder(fp: str):
    get_storage_manager().assert_fp_exists(fp)


def get_fns(fp: str):
    return get_storage_manager().get_fns(fp)


def llm_load(fp: str, **kwargs):
    return get_storage_manager().load(fp, **kwargs)


def llm_save(save_path: str, saved_obj: Any, **kwargs):
    get_storage_manager().save(save_path, to_save_obj=saved_obj, **kwargs)
