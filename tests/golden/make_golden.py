#!/usr/bin/env python3
"""Regenerate tests/golden/ from the reference checkout (run in the build container only).

The reference ships no tests; its data files are the only known answers for this
path.  This script copies those DATA files (no source code)
and derives the SRDF collision-pair table the validity check needs:

  config_database/*.dat   <- rrt_connect_planner/config_database/*.dat   (sampler / goal outputs)
  solutions/*.dat         <- rrt_connect_planner/solutions/*.dat         (planner outputs)
  polygon_log.txt         <- rrt_connect_planner_example/src/polygon_log.txt (logged support hull)
  srdf_pairs.json         <- derived from rrt_connect_planner/description/nao_robot_rfoot.srdf:168-324

/root/reference does not exist on the GPU box; nothing at test time reads it.
"""
import json
import re
import shutil
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent

# The 23 links that carry collision geometry: exactly the links named in the SRDF's
# "never in collision" block (SRDF:199-324) - SURVEY.md Appendix A.2.
GEOM_LINKS = ["torso", "HeadYaw_link", "HeadPitch_link"]
for s in "LR":
    GEOM_LINKS += [f"{s}{n}_link" for n in (
        "HipYawPitch", "HipRoll", "HipPitch", "KneePitch", "AnklePitch", "AnkleRoll",
        "ShoulderRoll", "ElbowRoll", "WristYaw")]
    GEOM_LINKS.append(f"{s.lower()}_gripper")

# Links moved by the joints of planning group "whole_body_rfoot" (SRDF:105-135):
# everything except torso / head (MoveIt group filter: a pair is checked iff >=1 link is in the group).
GROUP_LINKS = [l for l in GEOM_LINKS if l not in ("torso", "HeadYaw_link", "HeadPitch_link")]


def main():
    for sub in ("config_database", "solutions"):
        (OUT / sub).mkdir(exist_ok=True)
        for f in sorted((REF / "rrt_connect_planner" / sub).glob("*.dat")):
            shutil.copyfile(f, OUT / sub / f.name)
    shutil.copyfile(REF / "rrt_connect_planner_example/src/polygon_log.txt", OUT / "polygon_log.txt")

    srdf = (REF / "rrt_connect_planner/description/nao_robot_rfoot.srdf").read_text()
    raw = re.findall(r'<disable_collisions\s+link1\s*=\s*"([^"]*)"\s+link2\s*=\s*"([^"]*)"', srdf)
    out = {}
    for trim in (True, False):
        disabled = set()
        for a, b in raw:
            if trim:  # srdfdom trims attribute values
                a, b = a.strip(), b.strip()
            disabled.add(frozenset((a, b)))
        enabled_all, enabled_group = [], []
        for i, a in enumerate(GEOM_LINKS):
            for b in GEOM_LINKS[i + 1:]:
                if frozenset((a, b)) in disabled:
                    continue
                enabled_all.append([a, b])
                if a in GROUP_LINKS or b in GROUP_LINKS:
                    enabled_group.append([a, b])
        out["trim" if trim else "notrim"] = {
            "n_disabled_entries": len(raw),
            "enabled_all": enabled_all,          # isPathValid: whole robot, no group filter
            "enabled_group": enabled_group,      # isFeasible: group whole_body_rfoot
        }
    out["geom_links"] = GEOM_LINKS
    (OUT / "srdf_pairs.json").write_text(json.dumps(out, indent=1) + "\n")
    for k in ("trim", "notrim"):
        print(k, len(out[k]["enabled_all"]), len(out[k]["enabled_group"]), out[k]["n_disabled_entries"])


if __name__ == "__main__":
    main()
