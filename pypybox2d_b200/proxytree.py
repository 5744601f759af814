"""The reference's broadphase tree TOPOLOGY, replayed on the host for the tree observables (SURVEY §8 row a19 and §8f #2:
`tree_height`, `tree_balance`, `tree_quality`; broadphase.py:137-150 over dynamictree.py:275-314).

The device finds pairs and answers queries without a tree (the pair set and the query answers do not depend on the tree's
shape, SURVEY §8a note), so nothing on the step path needs this.  The observables do: they are functions of the order in
which leaves were inserted, removed and rotated.  `ProxyTree` keeps that shape in flat integer / float columns indexed by
pool slot -- the same slot numbering as the reference's node pool (first free slot first, a freed slot is the next one
handed out, the pool doubles when empty: dynamictree.py:92-113, 371-408), because `tree_quality` sums node perimeters in
pool order (dynamictree.py:299-314) and a float sum depends on its order.

What drives it: the facade (`World`) creates a leaf when a fixture gets its proxy (fixture.py:250-261), removes it when the
proxy is destroyed (fixture.py:263-267), and after every step re-inserts -- in the reference's synchronisation order, the
world's body list x fixtures x children, world.py:582-588 -- every proxy whose fat box the device replaced
(`DynamicTree.move_proxy`, dynamictree.py:152-188: leaf out, new fat box, leaf in).  Exact as long as a proxy is re-inserted at
most once per step, i.e. always without continuous collision; a time-of-impact sub-step that moves a proxy a second time
inside one step (world.py:790-800) is replayed as one move.
"""

NIL = -1


class ProxyTree(object):
    def __init__(self, starting_capacity=16):
        n = int(starting_capacity)
        self.parent = list(range(1, n)) + [NIL]     # doubles as the free list's `next` (dynamictree.py:68-74)
        self.child1 = [NIL] * n
        self.child2 = [NIL] * n
        self.height = [-1] * n                      # leaf 0, free slot -1
        self.lx = [0.0] * n
        self.ly = [0.0] * n
        self.ux = [0.0] * n
        self.uy = [0.0] * n
        self.root = NIL
        self.free = 0
        self.count = 0
        self.insertions = 0

    # -- node pool (dynamictree.py:371-408) --------------------------------------------------------------------------------
    def _take(self):
        if self.free == NIL:
            old = len(self.parent)
            self.parent.extend(list(range(old + 1, 2 * old)) + [NIL])
            for col in (self.child1, self.child2):
                col.extend([NIL] * old)
            self.height.extend([-1] * old)
            for col in (self.lx, self.ly, self.ux, self.uy):
                col.extend([0.0] * old)
            self.free = old
        i = self.free
        self.free = self.parent[i]
        self.parent[i] = self.child1[i] = self.child2[i] = NIL
        self.height[i] = 0
        self.count += 1
        return i

    def _give(self, i):
        self.parent[i] = self.free
        self.height[i] = -1
        self.free = i
        self.count -= 1

    # -- boxes ---------------------------------------------------------------------------------------------------------
    def _box(self, i):
        return (self.lx[i], self.ly[i], self.ux[i], self.uy[i])

    def _set_union(self, i, a, b):
        self.lx[i] = min(self.lx[a], self.lx[b])
        self.ly[i] = min(self.ly[a], self.ly[b])
        self.ux[i] = max(self.ux[a], self.ux[b])
        self.uy[i] = max(self.uy[a], self.uy[b])

    def _perimeter(self, i):                         # common.py:851-856
        return 2.0 * ((self.ux[i] - self.lx[i]) + (self.uy[i] - self.ly[i]))

    def _perimeter_with(self, i, box):
        lx, ly, ux, uy = box
        return 2.0 * ((max(self.ux[i], ux) - min(self.lx[i], lx)) + (max(self.uy[i], uy) - min(self.ly[i], ly)))

    # -- public: leaves ------------------------------------------------------------------------------------------------
    def create_leaf(self, fat_box):
        """A new leaf with the given (already fattened) box; returns its slot."""
        i = self._take()
        self.lx[i], self.ly[i], self.ux[i], self.uy[i] = (float(v) for v in fat_box)
        self._insert(i)
        return i

    def destroy_leaf(self, i):
        self._remove(i)
        self._give(i)

    def reinsert_leaf(self, i, fat_box):
        """`move_proxy` after the containment test failed: the new fat box is the caller's (the device computed it)."""
        self._remove(i)
        self.lx[i], self.ly[i], self.ux[i], self.uy[i] = (float(v) for v in fat_box)
        self._insert(i)

    def move_leaf(self, i, tight_box, displacement, extension, multiplier):
        """The whole of `DynamicTree.move_proxy` on the host (dynamictree.py:152-188); True when the leaf was re-inserted."""
        lx, ly, ux, uy = tight_box
        if self.lx[i] <= lx and self.ly[i] <= ly and ux <= self.ux[i] and uy <= self.uy[i]:
            return False
        lx, ly, ux, uy = lx - extension, ly - extension, ux + extension, uy + extension
        dx, dy = multiplier * displacement[0], multiplier * displacement[1]
        if dx < 0.0:
            lx = lx + dx
        else:
            ux = ux + dx
        if dy < 0.0:
            ly = ly + dy
        else:
            uy = uy + dy
        self.reinsert_leaf(i, (lx, ly, ux, uy))
        return True

    # -- public: observables (dynamictree.py:275-314) ------------------------------------------------------------------------
    @property
    def tree_height(self):
        return 0 if self.root == NIL else self.height[self.root]

    @property
    def tree_balance(self):
        # the reference's `max_balance` computes the value and falls off the end without a `return` (dynamictree.py:285-297)
        return None

    @property
    def max_balance(self):
        """What `max_balance` computes before it forgets to return it: the largest height difference between siblings."""
        best = 0
        for i, h in enumerate(self.height):
            if h > 1:
                best = max(best, abs(self.height[self.child2[i]] - self.height[self.child1[i]]))
        return best

    @property
    def tree_quality(self):
        if self.root == NIL:
            return 0.0
        total = sum([self._perimeter(i) for i, h in enumerate(self.height) if h >= 0])
        return total / self._perimeter(self.root)

    def structure(self):
        """Nested tuples of leaf slots, for tests: (left, right) per inner node."""
        def walk(i):
            return i if self.child1[i] == NIL else (walk(self.child1[i]), walk(self.child2[i]))
        return None if self.root == NIL else walk(self.root)

    # -- insertion (dynamictree.py:410-501) --------------------------------------------------------------------------------
    def _insert(self, leaf):
        self.insertions += 1
        if self.root == NIL:
            self.root = leaf
            self.parent[leaf] = NIL
            return
        box = self._box(leaf)
        node = self.root
        while self.child1[node] != NIL:
            area = self._perimeter(node)
            combined = self._perimeter_with(node, box)
            cost = 2.0 * combined                    # a new parent for this node and the leaf
            inherit = 2.0 * (combined - area)        # what every descent pays at least
            costs = []
            for ch in (self.child1[node], self.child2[node]):
                if self.child1[ch] == NIL:
                    costs.append(self._perimeter_with(ch, box) + inherit)
                else:
                    costs.append((self._perimeter_with(ch, box) - self._perimeter(ch)) + inherit)
            if cost < costs[0] and cost < costs[1]:
                break
            node = self.child1[node] if costs[0] < costs[1] else self.child2[node]
        sibling = node
        old_parent = self.parent[sibling]
        new_parent = self._take()
        self.parent[new_parent] = old_parent
        self._set_union(new_parent, leaf, sibling)
        self.height[new_parent] = self.height[sibling] + 1
        if old_parent != NIL:
            if self.child1[old_parent] == sibling:
                self.child1[old_parent] = new_parent
            else:
                self.child2[old_parent] = new_parent
        else:
            self.root = new_parent
        self.child1[new_parent] = sibling
        self.child2[new_parent] = leaf
        self.parent[sibling] = new_parent
        self.parent[leaf] = new_parent
        self._refit_upwards(new_parent)

    def _refit_upwards(self, node):
        while node != NIL:
            node = self._rotate(node)
            a, b = self.child1[node], self.child2[node]
            self.height[node] = 1 + max(self.height[a], self.height[b])
            self._set_union(node, a, b)
            node = self.parent[node]

    # -- removal (dynamictree.py:503-541) ----------------------------------------------------------------------------------
    def _remove(self, leaf):
        if leaf == self.root:
            self.root = NIL
            return
        parent = self.parent[leaf]
        sibling = self.child2[parent] if self.child1[parent] == leaf else self.child1[parent]
        grand = self.parent[parent]
        if grand != NIL:
            if self.child1[grand] == parent:
                self.child1[grand] = sibling
            else:
                self.child2[grand] = sibling
            self.parent[sibling] = grand
            self._give(parent)
            self._refit_upwards(grand)
        else:
            self.root = sibling
            self.parent[sibling] = NIL
            self._give(parent)

    # -- rotation (dynamictree.py:543-642) -------------------------------------------------------------------------------
    def _rotate(self, a):
        """Lift the taller child of `a` when the children's heights differ by more than one; returns the subtree's new top."""
        if self.child1[a] == NIL or self.height[a] < 2:
            return a
        b, c = self.child1[a], self.child2[a]
        balance = self.height[c] - self.height[b]
        if balance > 1:
            return self._lift(a, up=c, stay=b, up_is_child2=True)
        if balance < -1:
            return self._lift(a, up=b, stay=c, up_is_child2=False)
        return a

    def _lift(self, a, up, stay, up_is_child2):
        p, q = self.child1[up], self.child2[up]      # the lifted node's children: the taller stays with it, the other goes to `a`
        self.child1[up] = a
        top = self.parent[a]
        self.parent[up] = top
        self.parent[a] = up
        if top != NIL:
            if self.child1[top] == a:
                self.child1[top] = up
            else:
                self.child2[top] = up
        else:
            self.root = up
        keep, hand = (p, q) if self.height[p] > self.height[q] else (q, p)
        self.child2[up] = keep
        if up_is_child2:
            self.child2[a] = hand
        else:
            self.child1[a] = hand
        self.parent[hand] = a
        self._set_union(a, stay, hand)
        self._set_union(up, a, keep)
        self.height[a] = 1 + max(self.height[stay], self.height[hand])
        self.height[up] = 1 + max(self.height[a], self.height[keep])
        return up
