"""Peer discovery and P2P islands of optixNVLink (SDK/optixNVLink/optixNVLink.cpp:1698-1825) on plain data — no GPU, no NVML."""
from optix_raytracer_b200 import topology as T


def test_islands_from_cuda_peer_matrix():
    # two NVLink pairs and a lone device
    can = {(0, 1), (1, 0), (2, 3), (3, 2)}
    peers = T.peers_from_cuda(5, lambda a, b: (a, b) in can)
    assert peers == [0b00010, 0b00001, 0b01000, 0b00100, 0]
    islands = T.compute_p2p_islands(peers)
    assert islands == [0b00011, 0b01100, 0b10000]
    assert T.format_islands(islands) == "P2P ISLANDS: {0,1} {2,3} {4}"
    assert T.island_of(islands, 3) == 0b01100 and T.island_of(islands, 4) == 0b10000


def test_peers_from_nvlink_links_bridge_and_switch():
    ids = ["00000000:17:00.0", "0000:65:00.0", "00000000:B3:00.0"]
    link = lambda remote, active=True, p2p=True, switch=False: {"p2p": p2p, "active": active, "remote": remote, "switch": switch}  # noqa: E731
    # bridge: 0 <-> 1 active, 1 -> 2 inactive, one link without P2P capability, one to a device that is not visible
    links = [[link("00000000:65:00.0"), link("00000000:B3:00.0", p2p=False)], [link("0000:17:00.0"), link("00000000:b3:00.0", active=False)],
             [link("00000000:ff:00.0")]]
    assert T.peers_from_links(ids, links, require_nvlink=True) == [0b010, 0b001, 0]
    assert T.peers_from_links(ids, links, require_nvlink=False) == [0b010, 0b101, 0]
    # NVSwitch: every link ends at a switch; the devices with an active link form one fabric
    sw = [[link("00000000:01:00.0", switch=True)] * 2, [link("00000000:02:00.0", switch=True)], [link("00000000:03:00.0", switch=True, active=False)]]
    peers = T.peers_from_links(ids, sw, require_nvlink=True)
    assert peers == [0b010, 0b001, 0]
    assert T.compute_p2p_islands(peers) == [0b011, 0b100]
    assert T.compute_p2p_islands(T.peers_from_links(ids, sw, require_nvlink=False)) == [0b111]


def test_texture_sharing_plan_follows_optixNVLink():
    """loadTexture / getIslandDeviceWithLowestTextureUsage (SDK/optixNVLink/optixNVLink.cpp:1501-1561): one copy per island on the device
    with the least texture memory so far (the first such device; the scan starts from device 0 with a strict comparison), views for the
    rest of the island; without sharing every device loads its own."""
    islands = [0b0011, 0b1100]
    owners, usage = T.plan_texture_sharing(islands, [4.0, 4.0, 2.0], 4)
    assert owners == [[0, 0, 2, 2], [1, 1, 3, 3], [0, 0, 2, 2]]
    assert usage == [6.0, 4.0, 6.0, 4.0]
    owners, usage = T.plan_texture_sharing([0b1111], [1.0] * 5, 4)
    assert [o[0] for o in owners] == [0, 1, 2, 3, 0] and usage == [2.0, 1.0, 1.0, 1.0]
    owners, usage = T.plan_texture_sharing(islands, [4.0, 4.0], 4, share=False)
    assert owners == [[0, 1, 2, 3]] * 2 and usage == [8.0] * 4
    # a lone device is its own island
    owners, usage = T.plan_texture_sharing([0b011, 0b100], [3.0], 3)
    assert owners == [[0, 0, 2]] and usage == [3.0, 0.0, 3.0]
