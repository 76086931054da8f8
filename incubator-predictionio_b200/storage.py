"""Host-side mirror of the event-data types the ALS templates touch.

Mirrors (reference paths relative to the repository root):
  Event        data/src/main/scala/org/apache/predictionio/data/storage/Event.scala:42-54
  DataMap      data/src/main/scala/org/apache/predictionio/data/storage/DataMap.scala
  BiMap        data/src/main/scala/org/apache/predictionio/data/storage/BiMap.scala:28-167
  PEventStore  data/src/main/scala/org/apache/predictionio/data/store/PEventStore.scala:59-119
  $set/$unset/$delete fold   data/.../storage/PEventAggregator.scala:196-209, LEventAggregator

The storage *engine* (JDBC/HBase/ES backends, Event Server) is out of scope (SURVEY 8 / section 2 rows
11-15): events live in one JSON-lines file per app in the `pio import` / `pio export` format
(tools/src/main/scala/org/apache/predictionio/tools/imprt/FileToEvents.scala:93-103) under
$PIO_EVENTDATA_DIR (default ./pio_eventdata).  `find` returns a list of Event (the RDD stand-in)
or, for the bulk path, numpy column arrays.
"""
from __future__ import annotations

import datetime as _dt
import json
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np


def _coerce(v, typ):
    """json4s-style extraction of a JSON value into a Python type: primitives, List[T], Set[T], Optional[T],
    dataclasses (case classes; absent Optional fields become None)."""
    import dataclasses
    import typing
    origin = typing.get_origin(typ)
    args = typing.get_args(typ)
    if typ is None or typ is typing.Any:
        return v
    if origin is typing.Union:                       # Optional[T]
        inner = [a for a in args if a is not type(None)]
        return None if v is None else _coerce(v, inner[0])
    if origin in (list, typing.List):
        return [_coerce(x, args[0]) if args else x for x in v]
    if origin in (set, typing.Set, frozenset):
        return {(_coerce(x, args[0]) if args else x) for x in v}
    if typ in (list, set):
        return typ(v)
    if dataclasses.is_dataclass(typ):
        hints = typing.get_type_hints(typ)
        kw = {}
        for f in dataclasses.fields(typ):
            if f.name in v and v[f.name] is not None:
                kw[f.name] = _coerce(v[f.name], hints[f.name])
            elif typing.get_origin(hints[f.name]) is typing.Union and type(None) in typing.get_args(hints[f.name]):
                kw[f.name] = None
            elif f.default is not dataclasses.MISSING or f.default_factory is not dataclasses.MISSING:  # type: ignore
                continue
            else:
                raise DataMapException(f"The field {f.name} is required.")
        return typ(**kw)
    if typ is bool:
        if not isinstance(v, bool):
            raise DataMapException(f"{v!r} is not a Boolean")
        return v
    return typ(v)


class DataMapException(KeyError):
    pass


class DataMap:
    """JSON property bag with typed accessors (get / getOpt / getOrElse)."""

    __slots__ = ("fields",)

    def __init__(self, fields: Optional[Dict[str, Any]] = None):
        self.fields = dict(fields or {})

    def require(self, name: str) -> None:
        if name not in self.fields:
            raise DataMapException(f"The field {name} is required.")

    def contains(self, name: str) -> bool:
        return name in self.fields

    def get(self, name: str, typ=None):
        self.require(name)
        v = self.fields[name]
        if v is None:
            raise DataMapException(f"The required field {name} cannot be null.")
        return _coerce(v, typ) if typ is not None else v

    def getOpt(self, name: str, typ=None):
        v = self.fields.get(name)
        if v is None:
            return None
        return _coerce(v, typ) if typ is not None else v

    def getOrElse(self, name: str, default, typ=None):
        v = self.getOpt(name, typ)
        return default if v is None else v

    def extract(self, cls):
        """DataMap.extract[T] (DataMap.scala): the whole property bag as a case class - here a dataclass whose
        Optional[...] fields may be absent."""
        return _coerce(self.fields, cls)

    @staticmethod
    def fromJson(text: str) -> "DataMap":
        """DataMap(jsonString) of the reference (DataMap.scala companion)."""
        return DataMap(json.loads(text))

    def __add__(self, other: "DataMap") -> "DataMap":  # ++
        d = dict(self.fields)
        d.update(other.fields)
        return DataMap(d)

    def __sub__(self, keys: Iterable[str]) -> "DataMap":  # --
        ks = set(keys)
        return DataMap({k: v for k, v in self.fields.items() if k not in ks})

    def keySet(self):
        return set(self.fields)

    def isEmpty(self) -> bool:
        return not self.fields

    def toJson(self) -> Dict[str, Any]:
        return dict(self.fields)

    def __eq__(self, o):
        return isinstance(o, DataMap) and self.fields == o.fields

    def __repr__(self):
        return f"DataMap({self.fields})"


class PropertyMap(DataMap):
    """DataMap + first/last update time (result of aggregateProperties)."""

    __slots__ = ("firstUpdated", "lastUpdated")

    def __init__(self, fields, firstUpdated, lastUpdated):
        super().__init__(fields)
        self.firstUpdated = firstUpdated
        self.lastUpdated = lastUpdated

    def __eq__(self, o):   # PropertyMap.scala: equal fields AND equal first/last update times
        if isinstance(o, PropertyMap):
            return self.fields == o.fields and self.firstUpdated == o.firstUpdated and self.lastUpdated == o.lastUpdated
        return False

    __hash__ = None

    def __repr__(self):
        return f"PropertyMap({self.fields}, {self.firstUpdated}, {self.lastUpdated})"


def _parse_time(s) -> _dt.datetime:
    if isinstance(s, _dt.datetime):
        return s
    if s is None:
        return _dt.datetime.now(_dt.timezone.utc)
    s = str(s)
    if s.endswith("Z"):
        s = s[:-1] + "+00:00"
    t = _dt.datetime.fromisoformat(s)
    if t.tzinfo is None:
        t = t.replace(tzinfo=_dt.timezone.utc)
    return t


@dataclass
class Event:
    event: str
    entityType: str
    entityId: str
    targetEntityType: Optional[str] = None
    targetEntityId: Optional[str] = None
    properties: DataMap = field(default_factory=DataMap)
    eventTime: _dt.datetime = field(default_factory=lambda: _dt.datetime.now(_dt.timezone.utc))
    eventId: Optional[str] = None

    @staticmethod
    def from_json(d: Dict[str, Any]) -> "Event":
        for req in ("event", "entityType", "entityId"):
            if req not in d:
                raise ValueError(f"field {req} is required")  # EventValidation (Event.scala:68-167)
        return Event(event=d["event"], entityType=d["entityType"], entityId=str(d["entityId"]),
                     targetEntityType=d.get("targetEntityType"),
                     targetEntityId=None if d.get("targetEntityId") is None else str(d["targetEntityId"]),
                     properties=DataMap(d.get("properties") or {}), eventTime=_parse_time(d.get("eventTime")),
                     eventId=d.get("eventId"))

    def to_json(self) -> Dict[str, Any]:
        d = {"event": self.event, "entityType": self.entityType, "entityId": self.entityId,
             "properties": self.properties.toJson(), "eventTime": self.eventTime.isoformat()}
        if self.targetEntityType is not None:
            d["targetEntityType"] = self.targetEntityType
        if self.targetEntityId is not None:
            d["targetEntityId"] = self.targetEntityId
        if self.eventId is not None:
            d["eventId"] = self.eventId
        return d


class BiMap:
    """Immutable bi-directional map; `inverse` requires unique values (BiMap.scala:28-39)."""

    def __init__(self, m: Dict, _inv: Optional["BiMap"] = None):
        self._m = dict(m)
        # `val inverse` is built eagerly in the reference (BiMap.scala:33-39): a map with duplicated values fails at
        # construction (require -> IllegalArgumentException; ValueError here), and inverse.inverse is this object
        if _inv is None:
            rev = {v: k for k, v in self._m.items()}
            if len(rev) != len(self._m):
                raise ValueError("Failed to create reversed map. Cannot have duplicated values.")
            _inv = BiMap.__new__(BiMap)
            _inv._m = rev
            _inv._i = self
        self._i = _inv

    @property
    def inverse(self) -> "BiMap":
        return self._i

    def get(self, k):
        return self._m.get(k)

    def getOrElse(self, k, default):
        return self._m.get(k, default)

    def contains(self, k) -> bool:
        return k in self._m

    def __contains__(self, k):
        return k in self._m

    def apply(self, k):
        return self._m[k]

    __call__ = apply
    __getitem__ = apply

    def toMap(self) -> Dict:
        return dict(self._m)

    def toSeq(self):
        return list(self._m.items())

    @property
    def size(self) -> int:
        return len(self._m)

    def __len__(self):
        return len(self._m)

    def take(self, n: int) -> "BiMap":
        return BiMap(dict(list(self._m.items())[:n]))

    @staticmethod
    def stringInt(keys: Iterable[str]) -> "BiMap":
        """keys.distinct -> index in first-occurrence order (BiMap.scala:116-128; the reference's
        `distinct.collect` order is unspecified, so results must be compared by string id)."""
        m: Dict[str, int] = {}
        for k in keys:
            if k not in m:
                m[k] = len(m)
        return BiMap(m)

    @staticmethod
    def stringLong(keys: Iterable[str]) -> "BiMap":
        return BiMap.stringInt(keys)

    @staticmethod
    def stringDouble(keys: Iterable[str]) -> "BiMap":
        b = BiMap.stringInt(keys)
        return BiMap({k: float(v) for k, v in b._m.items()})


# --------------------------------------------------------------------------------------------------
# event store (file backed)
# --------------------------------------------------------------------------------------------------
def _data_dir() -> Path:
    return Path(os.environ.get("PIO_EVENTDATA_DIR", "pio_eventdata"))


def app_file(appName: str, channelName: Optional[str] = None) -> Path:
    name = appName if channelName is None else f"{appName}.{channelName}"
    return _data_dir() / f"{name}.jsonl"


def import_events(appName: str, events: Iterable[Event | Dict[str, Any]], channelName: Optional[str] = None) -> int:
    """`pio import`: append events to the app's JSON-lines file."""
    p = app_file(appName, channelName)
    p.parent.mkdir(parents=True, exist_ok=True)
    n = 0
    with open(p, "a") as f:
        for e in events:
            d = e.to_json() if isinstance(e, Event) else Event.from_json(e).to_json()
            f.write(json.dumps(d) + "\n")
            n += 1
    return n


def delete_app_data(appName: str, channelName: Optional[str] = None) -> None:
    p = app_file(appName, channelName)
    if p.exists():
        p.unlink()


def _iter_events(appName: str, channelName: Optional[str]) -> Iterator[Event]:
    p = app_file(appName, channelName)
    if not p.exists():
        raise FileNotFoundError(f"Invalid app name {appName}: no event data at {p}")  # Common.appNameToId
    with open(p) as f:
        for line in f:
            line = line.strip()
            if line:
                yield Event.from_json(json.loads(line))


_UNSET = object()


class LEventAggregator:
    """Fold of $set / $unset / $delete events into per-entity properties, restated from
    data/src/main/scala/org/apache/predictionio/data/storage/LEventAggregator.scala:42-146: events are sorted by event
    time per entity; $set merges (or creates), $unset removes keys (no-op without state), $delete drops the state;
    firstUpdated / lastUpdated are the earliest / latest time over ALL three event kinds (a $delete does not reset them)."""

    eventNames = ["$set", "$unset", "$delete"]

    @staticmethod
    def _fold(events: Iterable[Event]):
        dm: Optional[Dict[str, Any]] = None
        first = last = None
        for e in sorted(events, key=lambda x: x.eventTime):
            if e.event not in LEventAggregator.eventNames:
                continue
            if e.event == "$set":
                dm = dict(e.properties.fields) if dm is None else {**dm, **e.properties.fields}
            elif e.event == "$unset":
                if dm is not None:
                    dm = {k: v for k, v in dm.items() if k not in e.properties.fields}
            else:
                dm = None
            first = e.eventTime if first is None or e.eventTime < first else first
            last = e.eventTime if last is None or e.eventTime > last else last
        return dm, first, last

    @staticmethod
    def aggregateProperties(events: Iterable[Event]) -> Dict[str, PropertyMap]:
        groups: Dict[str, List[Event]] = {}
        for e in events:
            groups.setdefault(e.entityId, []).append(e)
        out = {}
        for k, evs in groups.items():
            dm, first, last = LEventAggregator._fold(evs)
            if dm is not None:
                out[k] = PropertyMap(dm, first, last)
        return out

    @staticmethod
    def aggregatePropertiesSingle(events: Iterable[Event]) -> Optional[PropertyMap]:
        dm, first, last = LEventAggregator._fold(events)
        return None if dm is None else PropertyMap(dm, first, last)


class PEventStore:
    """PEventStore.find / aggregateProperties over the file store (the `sc` argument is accepted and ignored)."""

    @staticmethod
    def find(appName: str, channelName: Optional[str] = None, startTime=None, untilTime=None,
             entityType: Optional[str] = None, entityId: Optional[str] = None,
             eventNames: Optional[Sequence[str]] = None, targetEntityType=_UNSET, targetEntityId=_UNSET,
             sc=None) -> List[Event]:
        """targetEntityType / targetEntityId follow Option[Option[String]]: omitted = no restriction,
        None = must be absent, "x" = must equal."""
        names = None if eventNames is None else set(eventNames)
        out = []
        for e in _iter_events(appName, channelName):
            if startTime is not None and e.eventTime < _parse_time(startTime):
                continue
            if untilTime is not None and e.eventTime >= _parse_time(untilTime):
                continue
            if entityType is not None and e.entityType != entityType:
                continue
            if entityId is not None and e.entityId != entityId:
                continue
            if names is not None and e.event not in names:
                continue
            if targetEntityType is not _UNSET and e.targetEntityType != targetEntityType:
                continue
            if targetEntityId is not _UNSET and e.targetEntityId != targetEntityId:
                continue
            out.append(e)
        return out

    @staticmethod
    def aggregateProperties(appName: str, entityType: str, channelName: Optional[str] = None, startTime=None,
                            untilTime=None, required: Optional[Sequence[str]] = None, sc=None
                            ) -> List[Tuple[str, PropertyMap]]:
        evs = PEventStore.find(appName, channelName, startTime, untilTime, entityType=entityType,
                               eventNames=["$set", "$unset", "$delete"])
        out = []
        for k, pm in LEventAggregator.aggregateProperties(evs).items():
            if required is not None and not all(r in pm.fields for r in required):
                continue
            out.append((k, pm))
        return out


class LEventStore:
    """Serving-time lookups (LEventStore.findByEntity, data/.../store/LEventStore.scala:76)."""

    @staticmethod
    def findByEntity(appName: str, entityType: str, entityId: str, channelName: Optional[str] = None,
                     eventNames: Optional[Sequence[str]] = None, targetEntityType=_UNSET, targetEntityId=_UNSET,
                     startTime=None, untilTime=None, limit: Optional[int] = None, latest: bool = True,
                     timeout=None) -> List[Event]:
        evs = PEventStore.find(appName, channelName, startTime, untilTime, entityType, entityId, eventNames,
                               targetEntityType, targetEntityId)
        evs.sort(key=lambda e: e.eventTime, reverse=latest)
        return evs if limit is None or limit < 0 else evs[:limit]
