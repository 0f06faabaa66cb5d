"""Which GPUs of the box can write each other's memory, and over what — the peer discovery of optixNVLink
(SDK/optixNVLink/optixNVLink.cpp:1698-1825: findPeersForDevice / findPeers / computeP2PIslands, :1617-1635 enablePeerAccess), restated
for a one-process-per-GPU launch.  Used by the shared result buffer (host.SharedResultBuffer): a rank may put the owner's buffer into
Params::result_buffer only when it sits in the owner's island.

peers[i] is a bit mask of the devices device i reaches (bit j = device j), as PerDeviceSampleState::peers; an island is `peers | 1 << i`.
With NVML (`pynvml`) every NVLink link of a device is examined like the reference does: P2P capability, link state when NVLink is
required, and the PCI bus id of the remote end, matched against the CUDA devices.  On an NVSwitch box (B200 HGX) the remote end of every
link is a switch, not a GPU — there the devices behind active links form one fabric and are all peers of each other.  (The reference's
own comparison `std::string(pci.busId) == pci_id` compares the remote id with itself, :1760-1769, so it marks every device a peer of
every device that has a usable link; on switch-less NVLink bridges that over-reports, here the ids are really compared.)  Without NVML
the CUDA peer-access matrix is used and every P2P connection is treated as NVLink, as the reference's #else branch does (:1781-1796)."""
from typing import Callable, List, Optional, Sequence

NVML_NVLINK_MAX_LINKS = 18
_CAP_P2P_SUPPORTED = 0
_REMOTE_TYPE_SWITCH = 2  # NVML_NVLINK_DEVICE_TYPE_SWITCH


def _norm_bus_id(s):
    if isinstance(s, bytes):
        s = s.decode()
    s = s.strip().lower()
    dom, _, rest = s.partition(":")
    return f"{int(dom, 16):08x}:{rest}" if rest else s


def peers_from_cuda(num_devices: int, can_access: Callable[[int, int], bool]) -> List[int]:
    """The #else branch: cudaDeviceCanAccessPeer for every ordered pair."""
    peers = [0] * num_devices
    for i in range(num_devices):
        for j in range(num_devices):
            if i != j and can_access(i, j):
                peers[i] |= 1 << j
    return peers


def peers_from_links(bus_ids: Sequence[str], links: Sequence[Sequence[dict]], require_nvlink: bool = True) -> List[int]:
    """The NVML branch on plain data: bus_ids[i] = PCI bus id of CUDA device i; links[i] = one dict per NVLink link of device i with keys
    p2p (capability NVML_NVLINK_CAP_P2P_SUPPORTED), active (nvmlDeviceGetNvLinkState), remote (bus id of the remote end) and switch (the
    remote end is an NVSwitch)."""
    ids = [_norm_bus_id(b) for b in bus_ids]
    n = len(ids)
    peers = [0] * n
    on_fabric = [False] * n
    for i in range(n):
        for ln in links[i]:
            if not ln.get("p2p"):
                continue
            if require_nvlink and not ln.get("active"):
                continue
            if ln.get("switch"):
                on_fabric[i] = True
                continue
            remote = _norm_bus_id(ln.get("remote", ""))
            for j in range(n):
                if j != i and ids[j] == remote:
                    peers[i] |= 1 << j
    fabric = sum(1 << i for i in range(n) if on_fabric[i])
    for i in range(n):
        if on_fabric[i]:
            peers[i] |= fabric & ~(1 << i)
    return peers


def compute_p2p_islands(peers: Sequence[int]) -> List[int]:
    """computeP2PIslands (:1698-1712): the distinct `peers | self` masks in device order."""
    islands: List[int] = []
    for i, p in enumerate(peers):
        isl = p | (1 << i)
        if isl not in islands:
            islands.append(isl)
    return islands


def island_of(islands: Sequence[int], device: int) -> int:
    for isl in islands:
        if isl >> device & 1:
            return isl
    return 1 << device


def plan_texture_sharing(islands: Sequence[int], texture_bytes: Sequence[float], num_devices: int, share: bool = True):
    """loadTextures / loadTexture / getIslandDeviceWithLowestTextureUsage (:1501-1590): where the copies of every texture go.
    With sharing, each P2P island keeps ONE copy of a texture, on the device of the island that holds the least texture memory so far
    (the first such device in index order), and every other device of the island samples that copy through its own texture object;
    without, every device loads its own.  Returns (owners, usage): owners[t][d] = the device whose array device d samples for texture t
    (d itself where it holds a copy), usage[d] = bytes of texture arrays on device d."""
    usage = [0.0] * num_devices
    owners = []
    for size in texture_bytes:
        own = [None] * num_devices
        if share:
            for isl in islands:
                best, best_usage = 0, float("inf")   # the reference starts from device 0 and takes a strictly lower usage only
                for d in range(isl.bit_length()):
                    if isl >> d & 1 and usage[d] < best_usage:
                        best, best_usage = d, usage[d]
                usage[best] += size
                for d in range(isl.bit_length()):
                    if isl >> d & 1:
                        own[d] = best
        else:
            for d in range(num_devices):
                usage[d] += size
                own[d] = d
        owners.append(own)
    return owners, usage


def format_islands(islands: Sequence[int]) -> str:
    """printIsland's text: "P2P ISLANDS: {0,1,2,3} {4}"."""
    return "P2P ISLANDS: " + " ".join("{" + ",".join(str(b) for b in range(isl.bit_length()) if isl >> b & 1) + "}" for isl in islands)


def find_peers(num_devices: Optional[int] = None, require_nvlink: bool = True) -> List[int]:
    """findPeers for the visible CUDA devices: NVML when it loads, else the CUDA peer matrix."""
    import torch
    n = torch.cuda.device_count() if num_devices is None else num_devices
    try:
        import pynvml as nv
        nv.nvmlInit()
    except Exception:  # noqa: BLE001 — "NVML NOT SUPPORTED. Cannot query nvlink. Treating all P2P connections as nvlink."
        return peers_from_cuda(n, torch.cuda.can_device_access_peer)
    try:
        bus_ids, links = [], []
        for i in range(n):
            bus = torch.cuda.get_device_properties(i)
            bus_id = f"{getattr(bus, 'pci_domain_id', 0):08x}:{getattr(bus, 'pci_bus_id', 0):02x}:{getattr(bus, 'pci_device_id', 0):02x}.0"
            h = nv.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
            bus_ids.append(bus_id)
            dev_links = []
            for link in range(NVML_NVLINK_MAX_LINKS):
                try:
                    cap = nv.nvmlDeviceGetNvLinkCapability(h, link, _CAP_P2P_SUPPORTED)
                except nv.NVMLError:
                    continue
                ln = {"p2p": bool(cap), "active": False, "remote": "", "switch": False}
                try:
                    ln["active"] = nv.nvmlDeviceGetNvLinkState(h, link) == nv.NVML_FEATURE_ENABLED
                    ln["remote"] = nv.nvmlDeviceGetNvLinkRemotePciInfo(h, link).busId
                    ln["switch"] = nv.nvmlDeviceGetNvLinkRemoteDeviceType(h, link) == _REMOTE_TYPE_SWITCH
                except (nv.NVMLError, AttributeError):
                    pass
                dev_links.append(ln)
            links.append(dev_links)
        peers = peers_from_links(bus_ids, links, require_nvlink)
        if not any(peers) and n > 1:
            # no NVLink information at all (virtualised NVML): fall back to what CUDA reports, like the build without NVML
            peers = peers_from_cuda(n, torch.cuda.can_device_access_peer)
        return peers
    except Exception:  # noqa: BLE001
        return peers_from_cuda(n, torch.cuda.can_device_access_peer)
    finally:
        try:
            nv.nvmlShutdown()
        except Exception:  # noqa: BLE001
            pass
